"""Class-API mirror of ``HumanoidDancingEnv`` (humanoid_dancing_env/dancing_env.py:23-154) on the CUDA engine.

Same constructor, ``reset``/``step`` signatures, spaces, ``metadata`` and ``info`` keys as the reference class; physics
(RK4, PGS), observation, reward, termination and the dance bookkeeping run in the fused kernel (a one-env batch).
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np

from ..spaces import _GymEnv
from ..vector_env import B200VectorEnv

DANCE_MOVES = {   # dancing_env.py:57-68
    "basic_step": {"difficulty": 1, "energy": 0.5, "style_points": 10},
    "spin": {"difficulty": 2, "energy": 1.0, "style_points": 20},
    "jump": {"difficulty": 2, "energy": 1.5, "style_points": 25},
    "moonwalk": {"difficulty": 3, "energy": 0.8, "style_points": 40},
    "robot_wave": {"difficulty": 2, "energy": 0.6, "style_points": 30},
    "freeze": {"difficulty": 1, "energy": 0.2, "style_points": 15},
    "hip_hop_bounce": {"difficulty": 2, "energy": 0.7, "style_points": 25},
    "breakdance_toprock": {"difficulty": 3, "energy": 1.2, "style_points": 35},
    "salsa_basic": {"difficulty": 2, "energy": 0.8, "style_points": 28},
    "ballet_pirouette": {"difficulty": 4, "energy": 1.0, "style_points": 50},
}


class HumanoidDancingEnv(_GymEnv):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 60}

    def __init__(self, render_mode: Optional[str] = None, **kwargs):
        if render_mode is not None:
            raise NotImplementedError("render_mode must be None: rendering is not part of the B200 engine")
        self.render_mode = render_mode
        self.dt = 0.01667; self.max_episode_steps = 3600
        self.floor_radius = 10.0; self.stage_height = 0.5
        self.bpm = 120; self.beat_interval = 60.0 / self.bpm
        self.dance_moves = DANCE_MOVES
        self._vec = B200VectorEnv("humanoid_dancing", 1, device=kwargs.get("device", 0), seed=kwargs.get("seed", 0) or 0)
        self.model = self._vec.tables; self.data = self._vec.batch
        self.num_joints = int(self.model.nu)
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
        # _generate_dance_sequence (dancing_env.py:896-905): 20 x (choice of 10 moves, uniform(1, 3)) from the env's Generator
        inject = np.zeros((1, 40), np.float32)
        for k in range(20):
            inject[0, 2 * k] = self.np_random.integers(0, 10); inject[0, 2 * k + 1] = self.np_random.uniform(1.0, 3.0)
        obs, _ = self._vec.reset(options={"inject": inject})
        st = self._state()
        info = {"episode_stats": st["episode_stats"], "current_move": self.dance_moves[list(self.dance_moves)[0]],
                "beat_phase": 0.0, "combo_multiplier": st["combo_multiplier"]}
        return obs[0].cpu().numpy(), info

    def step(self, action: np.ndarray):
        a = np.asarray(action, np.float32).reshape(1, -1)
        obs, rew, term, trunc, infos = self._vec.step(a)
        done = bool(term[0]) or bool(trunc[0])
        o = (infos["final_obs"][0] if done else obs[0]).cpu().numpy()
        st = self._state()
        names = list(self.dance_moves)
        info = {"episode_stats": st["episode_stats"], "current_move": self.dance_moves[names[st["current_move_idx"] % len(names)]],
                "beat_phase": st["time_since_last_beat"] / self.beat_interval, "combo_multiplier": st["combo_multiplier"],
                "crowd_excitement": st["crowd_excitement"], "performance_score": st["performance_score"]}
        return o, float(rew[0]), bool(term[0]), bool(trunc[0]), info

    def _state(self) -> Dict[str, Any]:
        ti, tf = self._vec.task_state()       # the finished episode's values on a terminal step
        ti = ti[0].cpu().numpy(); tf = tf[0].cpu().numpy()
        d = tf[2:8].view(np.float64)       # time_since_last_beat, combo_multiplier, move_start_time
        stats = {"total_score": float(tf[0]), "perfect_moves": 0, "good_moves": 0, "missed_beats": 0,
                 "longest_combo": int(ti[6]), "energy_used": float(tf[11]), "time_on_beat": float(tf[12]),
                 "creativity_score": 0.0, "crowd_rating": float(tf[1])}
        return {"episode_stats": stats, "current_move_idx": int(ti[3]), "time_since_last_beat": float(d[0]),
                "combo_multiplier": float(d[1]), "crowd_excitement": float(tf[1]), "performance_score": float(tf[0]),
                "current_step": int(ti[0])}

    def render(self):
        return None

    def close(self):
        self._vec.close()
