"""Class-API mirror of ``HumanoidConstructionEnv`` (humanoid_construction_env/construction_env.py:24-153) on the CUDA
engine.  Same constructor, keyword-only ``reset`` (:547), ``step``, ``metadata`` and ``info`` keys; the observation space
declares the 135 entries the reference actually returns (it declares 125).
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np

from ..spaces import _GymEnv
from ..vector_env import B200VectorEnv

TASK_TYPES = ["stack_blocks", "operate_crane", "transport_material", "build_structure"]


class HumanoidConstructionEnv(_GymEnv):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 50}

    def __init__(self, render_mode: Optional[str] = None, **kwargs):
        if render_mode is not None:
            raise NotImplementedError("render_mode must be None: rendering is not part of the B200 engine")
        self.render_mode = render_mode
        self.dt = 0.02; self.max_episode_steps = 3000; self.task_types = TASK_TYPES; self.max_blocks = 20
        self._vec = B200VectorEnv("humanoid_construction", 1, device=kwargs.get("device", 0), seed=kwargs.get("seed", 0) or 0)
        self.model = self._vec.tables; self.data = self._vec.batch
        self.num_joints = int(self.model.nu)
        self.action_space = self._vec.single_action_space
        self.observation_space = self._vec.single_observation_space
        self.np_random = np.random.default_rng(kwargs.get("seed"))

    def reset(self, *, seed: Optional[int] = None, options: Optional[Dict[str, Any]] = None) -> Tuple[np.ndarray, Dict[str, Any]]:
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        r = self.np_random
        inject = np.array([[r.integers(0, 4), r.uniform(0, 5), r.uniform(0, 0.5), r.uniform(15, 35)]], np.float32)   # :558-572
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
        return {"task": TASK_TYPES[int(ti[1])], "task_progress": float(tf[1]), "blocks_placed": 0, "safety_violations": 0,
                "episode_stats": {"blocks_placed": 0, "materials_transported": 0, "crane_operations": 0, "safety_violations": 0,
                                  "tasks_completed": int(ti[3]), "total_reward": float(tf[0])},
                "weather": {"wind": float(tf[2]), "rain": float(tf[3]), "temperature": float(tf[4])}}

    def render(self):
        return None

    def close(self):
        self._vec.close()
