"""Class-API mirror of ``QuadrupedParkourEnv`` (quadruped_parkour_env/parkour_env.py:20-82) on the CUDA engine.

Same constructor, ``reset``/``step`` signatures, spaces, ``metadata`` and ``info`` keys as the reference class; the
physics, observation, reward and termination run in the fused kernel (a one-env batch).  ``render_mode`` other than
``None`` is rejected: rendering is outside the hot path.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np

from ..spaces import _GymEnv
from ..tasks import TASKS
from ..vector_env import B200VectorEnv


class QuadrupedParkourEnv(_GymEnv):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 100}

    def __init__(self, render_mode: Optional[str] = None, **kwargs):
        if render_mode is not None:
            raise NotImplementedError("render_mode must be None: rendering is not part of the B200 engine")
        self.render_mode = render_mode
        self.dt = 0.01; self.frame_skip = 10; self.max_episode_steps = 6000
        self.course_length = 100.0; self.course_width = 20.0
        self.start_pos = np.array([2.0, 0.0, 0.6]); self.finish_pos = np.array([98.0, 0.0, 0.0])
        self._vec = B200VectorEnv("quadruped_parkour", 1, device=kwargs.get("device", 0), seed=kwargs.get("seed", 0) or 0)
        self.model = self._vec.tables          # compiled tables stand in for mujoco.MjModel
        self.data = self._vec.batch            # device state stands in for mujoco.MjData
        self.action_space = self._vec.single_action_space
        self.observation_space = self._vec.single_observation_space
        self.np_random = None
        self.seed(kwargs.get("seed"))

    def seed(self, seed: Optional[int] = None):
        self.np_random = np.random.default_rng(seed)
        return [seed]

    def reset(self, seed: Optional[int] = None, options: Optional[Dict] = None) -> Tuple[np.ndarray, Dict]:
        if seed is not None:
            self.seed(seed)
        # the two draws of _randomize_obstacles (parkour_env.py:757-774) come from the env's numpy Generator
        inject = np.zeros((1, 4), np.float32)
        inject[0, 0] = self.np_random.uniform(-1.5, 1.5); inject[0, 1] = self.np_random.uniform(-1.0, 1.0)
        obs, _ = self._vec.reset(options={"inject": inject})
        return obs[0].cpu().numpy(), self._info()

    def step(self, action: np.ndarray):
        a = np.asarray(action, np.float32).reshape(1, -1)
        # the class API has no auto-reset: snapshot the terminal observation the kernel keeps in final_obs
        obs, rew, term, trunc, infos = self._vec.step(a)
        done = bool(term[0]) or bool(trunc[0])
        o = (infos["final_obs"][0] if done else obs[0]).cpu().numpy()
        return o, float(rew[0]), bool(term[0]), bool(trunc[0]), self._info(prev_episode=done)

    def _info(self, prev_episode: bool = False) -> Dict[str, Any]:
        ti, tf, xpos = self._vec.task_state(with_xpos=True)   # the finished episode's values on a terminal step
        ti = ti[0].cpu().numpy(); tf = tf[0].cpu().numpy()
        x = float(xpos[0, self.model.name2id("body", "torso"), 0])
        return {"step_count": int(ti[0]), "episode_reward": float(tf[0]), "max_forward_progress": float(tf[2]),
                "checkpoints_reached": int(bin(int(ti[1])).count("1")), "fall_count": int(ti[2]),
                "course_completion": min(1.0, max(0.0, (x - 2.0) / 96.0))}

    def render(self):
        return None

    def close(self):
        self._vec.close()
