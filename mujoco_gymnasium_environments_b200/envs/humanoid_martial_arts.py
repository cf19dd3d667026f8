"""Class-API mirror of ``HumanoidMartialArtsEnv`` (humanoid_martial_arts_env/martial_arts_env.py:33-132) on the CUDA
engine.  Same constructor, ``reset(seed, options)`` (:442), ``step``, ``metadata`` and ``info`` keys (:623-630); the
observation space declares the 113 entries the reference actually returns (it declares 85).
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np

from ..spaces import _GymEnv
from ..vector_env import B200VectorEnv


class HumanoidMartialArtsEnv(_GymEnv):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 60}

    def __init__(self, render_mode: Optional[str] = None, **kwargs):
        if render_mode is not None:
            raise NotImplementedError("render_mode must be None: rendering is not part of the B200 engine")
        self.render_mode = render_mode
        self.dt = 0.01667; self.max_episode_steps = 6000; self.robot_height = 1.75
        self._vec = B200VectorEnv("humanoid_martial_arts", 1, device=kwargs.get("device", 0), seed=kwargs.get("seed", 0) or 0)
        self.model = self._vec.tables; self.data = self._vec.batch
        self.num_joints = int(self.model.nu)
        self.action_space = self._vec.single_action_space
        self.observation_space = self._vec.single_observation_space
        self.np_random = np.random.default_rng(kwargs.get("seed"))

    def reset(self, seed: Optional[int] = None, options: Optional[Dict[str, Any]] = None) -> Tuple[np.ndarray, Dict[str, Any]]:
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        r = self.np_random
        inject = np.array([[r.uniform(-0.5, 0.5), r.uniform(-0.5, 0.5)]], np.float32)      # :461-463
        obs, _ = self._vec.reset(options={"inject": inject})
        return obs[0].cpu().numpy(), self._info()

    def step(self, action: np.ndarray):
        a = np.asarray(action, np.float32).reshape(1, -1)
        obs, rew, term, trunc, infos = self._vec.step(a)
        done = bool(term[0]) or bool(trunc[0])
        o = (infos["final_obs"][0] if done else obs[0]).cpu().numpy()
        return o, float(rew[0]), bool(term[0]), bool(trunc[0]), self._info()

    def _info(self) -> Dict[str, Any]:
        ti, tf = self._vec.task_state()       # the finished episode's values on a terminal step
        ti = ti[0].cpu().numpy(); tf = tf[0].cpu().numpy()
        return {"episode_stats": {"techniques_performed": int(ti[1]), "successful_combos": 0, "balance_maintained": 0,
                                  "max_power_generated": 0.0, "total_distance_moved": 0.0, "falls": int(ti[3])},
                "combo_chain": [], "stance_stability": float(tf[1]), "current_step": int(ti[0])}

    def render(self):
        return None

    def close(self):
        self._vec.close()
