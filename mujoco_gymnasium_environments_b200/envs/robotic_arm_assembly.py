"""Class-API mirror of ``RoboticArmAssemblyEnv`` (robotic_arm_assembly_env/assembly_env.py:18-95) on the CUDA engine.
Same constructor (``render_mode``, ``config`` stored and never read, :28-33), ``reset(seed, options)`` (:162), ``step``,
``metadata`` and ``info`` keys (:474-484).  Like the reference, the constructor ends with ``reset()`` (:95).
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np

from ..spaces import _GymEnv
from ..tasks import ARM_SEQUENCE
from ..vector_env import B200VectorEnv

_STATUS = ["in_bin", "held", "assembled", "dropped"]
_PHASE = {0: "idle", 1: "pickup", 2: "transport", 3: "align", 4: "insert"}


class RoboticArmAssemblyEnv(_GymEnv):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 50}

    def __init__(self, render_mode: Optional[str] = None, config: Optional[Dict] = None, **kwargs):
        if render_mode is not None:
            raise NotImplementedError("render_mode must be None: rendering is not part of the B200 engine")
        self.render_mode = render_mode
        self.config = config or {}
        self.max_episode_steps = 150000; self.control_frequency = 50; self.simulation_frequency = 500; self.skip_frames = 10
        self.assembly_tolerance = 0.002; self.force_threshold = 50.0; self.gentle_force_threshold = 10.0
        self.assembly_sequence = list(ARM_SEQUENCE)
        self._vec = B200VectorEnv("robotic_arm_assembly", 1, device=kwargs.get("device", 0), seed=kwargs.get("seed", 0) or 0)
        self.model = self._vec.tables; self.data = self._vec.batch
        self.action_space = self._vec.single_action_space
        self.observation_space = self._vec.single_observation_space
        self.reset()

    def reset(self, seed: Optional[int] = None, options: Optional[Dict[str, Any]] = None) -> Tuple[np.ndarray, Dict[str, Any]]:
        obs, _ = self._vec.reset()           # the reference's reset draws nothing (:162-192)
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
        return {"step_count": int(ti[0]),
                "assembly_progress": {c: bool((int(ti[1]) >> k) & 1) for k, c in enumerate(ARM_SEQUENCE)},
                "component_status": {c: _STATUS[int(ti[5 + k])] for k, c in enumerate(ARM_SEQUENCE)},
                "task_phase": _PHASE[int(ti[4])], "held_component": ARM_SEQUENCE[int(ti[3])] if int(ti[3]) >= 0 else None,
                "cumulative_reward": float(tf[0]), "success": (int(ti[1]) & 0x1ff) == 0x1ff}

    def render(self):
        return None

    def close(self):
        self._vec.close()
