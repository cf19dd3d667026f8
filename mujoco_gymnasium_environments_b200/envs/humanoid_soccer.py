"""Class-API mirror of ``HumanoidSoccerEnv`` (humanoid_soccer_env/soccer_env.py:24-118) on the CUDA engine.

Same constructor, ``reset``/``step`` signatures, spaces, ``metadata`` and ``info`` keys as the reference class; physics
(Euler, PGS), goalkeeper / wind forces, observation, reward and termination run in the fused kernel (a one-env batch).
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np

from ..spaces import _GymEnv
from ..vector_env import B200VectorEnv


class HumanoidSoccerEnv(_GymEnv):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 50}

    def __init__(self, render_mode: Optional[str] = None, **kwargs):
        if render_mode is not None:
            raise NotImplementedError("render_mode must be None: rendering is not part of the B200 engine")
        self.render_mode = render_mode
        self.dt = 0.02; self.max_episode_steps = 5000          # the class value, not the registered 2500
        self.field_length = 50.0; self.field_width = 30.0; self.goal_width = 7.32; self.goal_height = 2.44
        self._vec = B200VectorEnv("humanoid_soccer", 1, device=kwargs.get("device", 0), seed=kwargs.get("seed", 0) or 0)
        self.model = self._vec.tables; self.data = self._vec.batch
        self.num_joints = int(self.model.nu)
        self.action_space = self._vec.single_action_space
        self.observation_space = self._vec.single_observation_space
        self._torso = self.model.name2id("body", "torso"); self._ball = self.model.name2id("body", "ball")
        self.np_random = None
        self.seed(kwargs.get("seed"))

    def seed(self, seed: Optional[int] = None):
        self.np_random = np.random.default_rng(seed)
        return [seed]

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None) -> Tuple[np.ndarray, dict]:
        if seed is not None:
            self.seed(seed)
        r = self.np_random
        # the draws of _randomize_initial_state / _update_environmental_factors in the reference's order (:454-508)
        inject = np.zeros((1, 36), np.float32)
        inject[0, 0] = r.uniform(-15.0, -5.0); inject[0, 1] = r.uniform(-10.0, 10.0); inject[0, 2] = r.uniform(-0.5, 0.5)
        inject[0, 3:32] = [r.uniform(-0.1, 0.1) for _ in range(29)]
        inject[0, 32] = r.uniform(-2.0, 2.0); inject[0, 33] = r.uniform(0.0, 2.0); inject[0, 34] = r.uniform(0, 2 * np.pi)
        inject[0, 35] = r.uniform(0.05, 0.15)
        obs, _ = self._vec.reset(options={"inject": inject})
        st = self._state()
        return obs[0].cpu().numpy(), {k: st[k] for k in ("episode_stats", "ball_position", "robot_position", "goal_distance")}

    def step(self, action: np.ndarray):
        a = np.asarray(action, np.float32).reshape(1, -1)
        obs, rew, term, trunc, infos = self._vec.step(a)
        done = bool(term[0]) or bool(trunc[0])
        o = (infos["final_obs"][0] if done else obs[0]).cpu().numpy()
        return o, float(rew[0]), bool(term[0]), bool(trunc[0]), self._state()

    def _state(self) -> Dict[str, Any]:
        ti, tf = self._vec.task_state()       # the finished episode's values on a terminal step
        ti = ti[0].cpu().numpy(); tf = tf[0].cpu().numpy()
        ball = tf[1:4].astype(np.float64); robot = tf[4:7].astype(np.float64)     # positions of the last forward pass
        quat_up = float(self._vec._obs[0, 50] ** 2 - self._vec._obs[0, 51] ** 2 - self._vec._obs[0, 52] ** 2 + self._vec._obs[0, 53] ** 2)
        stats = {"goals_scored": int(ti[2]), "ball_contacts": int(ti[3]), "distance_traveled": float(tf[12]),
                 "time_upright": float(tf[11]), "max_ball_speed": float(tf[13])}
        return {"episode_stats": stats, "ball_position": ball, "robot_position": robot,
                "goal_distance": float(np.linalg.norm(robot - np.array([24.5, 0.0, 0.0]))),
                "ball_contact": bool(ti[5]) if len(ti) > 5 else False, "robot_upright": quat_up > 0.7,
                "goal_scored": bool(ti[1])}

    def render(self):
        return None

    def close(self):
        self._vec.close()
