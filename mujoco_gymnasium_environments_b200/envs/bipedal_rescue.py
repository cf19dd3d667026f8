"""Class-API mirror of ``BipedalRescueEnv`` (bipedal_rescue_env/rescue_env.py:26-119) on the CUDA engine.

Same constructor, ``reset``/``step`` signatures, spaces, ``metadata`` and ``info`` keys as the reference class; physics
(RK4, PGS), victim pickup / drop-off, observation, reward and termination run in the fused kernel (a one-env batch).
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np

from ..spaces import _GymEnv
from ..vector_env import B200VectorEnv


class BipedalRescueEnv(_GymEnv):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 50}

    def __init__(self, render_mode: Optional[str] = None, **kwargs):
        if render_mode is not None:
            raise NotImplementedError("render_mode must be None: rendering is not part of the B200 engine")
        self.render_mode = render_mode
        self.dt = 0.02; self.max_episode_steps = 10000
        self.num_victims = 5; self.carry_capacity = 2; self.energy_limit = 1000.0
        self.safe_zone_pos = np.array([20.0, 0.0, 0.0]); self.safe_zone_radius = 3.0
        self.victim_priorities = [0.8, 1.0, 0.7, 0.9, 1.0]      # rescue_env.py:58
        self._vec = B200VectorEnv("bipedal_rescue", 1, device=kwargs.get("device", 0), seed=kwargs.get("seed", 0) or 0)
        self.model = self._vec.tables; self.data = self._vec.batch
        self.action_space = self._vec.single_action_space
        self.observation_space = self._vec.single_observation_space
        self.np_random = None
        self.seed(kwargs.get("seed"))

    def seed(self, seed: Optional[int] = None):
        self.np_random = np.random.default_rng(seed)
        return [seed]

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None) -> Tuple[np.ndarray, dict]:
        if seed is not None:
            self.seed(seed)
        r = self.np_random
        inject = np.zeros((1, 12), np.float32)        # _randomize_initial_state draws in the reference's order (:473-508)
        inject[0, 0] = r.uniform(-5.0, 5.0); inject[0, 1] = r.uniform(-5.0, 5.0)
        for k in range(5):
            inject[0, 2 + 2 * k] = r.uniform(-1.0, 1.0); inject[0, 3 + 2 * k] = r.uniform(-1.0, 1.0)
        obs, _ = self._vec.reset(options={"inject": inject})
        o = obs[0].cpu().numpy()
        st = self._state(o)
        return o, {"episode_stats": st["episode_stats"], "robot_position": st["robot_position"],
                   "victims_remaining": st["victims_remaining"], "energy_remaining": st["energy_remaining"]}

    def step(self, action: np.ndarray):
        a = np.asarray(action, np.float32).reshape(1, -1)
        obs, rew, term, trunc, infos = self._vec.step(a)
        done = bool(term[0]) or bool(trunc[0])
        o = (infos["final_obs"][0] if done else obs[0]).cpu().numpy()
        return o, float(rew[0]), bool(term[0]), bool(trunc[0]), self._state(o)

    def _state(self, o) -> Dict[str, Any]:
        ti, tf = self._vec.task_state()       # the finished episode's values on a terminal step
        ti = ti[0].cpu().numpy(); tf = tf[0].cpu().numpy()
        q = o[55:59]
        stats = {"victims_rescued": int(ti[11]), "distance_traveled": float(tf[6]), "energy_used": float(tf[7]),
                 "time_to_first_rescue": None if tf[8] < 0 else float(tf[8]), "falls": int(ti[9]), "collisions": int(ti[10])}
        return {"episode_stats": stats, "robot_position": o[52:55].astype(np.float64), "victims_remaining": 5 - int(ti[11]),
                "victims_carried": int(bin(int(ti[2])).count("1")), "energy_remaining": float(tf[1]),
                "robot_upright": bool(q[0] ** 2 - q[1] ** 2 - q[2] ** 2 + q[3] ** 2 > 0.7)}

    # rescue_env.py:59-60 keeps two Python lists; bipedal_rescue_env/test_rescue.py:298-299 tests membership in them.  The kernel keeps
    # them as bit masks (task state ti[1] rescued, ti[2] carried), so the lists come back in victim-index order, not pick-up order
    @property
    def victims_rescued(self):
        return self._victim_list(1)

    @property
    def victims_carried(self):
        return self._victim_list(2)

    def _victim_list(self, col: int):
        ti, _ = self._vec.task_state()
        bits = int(ti[0, col])
        return [i for i in range(self.num_victims) if (bits >> i) & 1]

    def render(self):
        return None

    def close(self):
        self._vec.close()
