"""CPU restatement of the reference task layers on top of the fp64 oracle.  TEST INFRASTRUCTURE ONLY.

Each class follows one reference env's order of operations including its
index-aliasing quirks, but drives ``oracle.ref`` (the C restatement of mj_step) instead of MuJoCo.
Used as (i) the checker for the fused CUDA task kernels and (ii) bench.py's CPU baseline ("port").
"""
from __future__ import annotations

import os

import numpy as np

from . import ref

_TABLES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mujoco_gymnasium_environments_b200", "tables")


def _load(task):
    from mujoco_gymnasium_environments_b200.mjcf import ModelTables
    return ModelTables.load(os.path.join(_TABLES, task + ".npz"))


class QuadrupedParkourRef:
    """quadruped_parkour_env/parkour_env.py restated: __init__ :34-82, reset :314-354, step :356-394,
    _get_observation :396-468, _calculate_reward :646-725, _is_terminated :727-755."""

    OBSTACLES = [(8.0, 1, 0.225, 0.3), (16.0, 2, 0.2, 0.6), (24.0, 3, 0.5, 0.8), (30.0, 4, 0.6, 0.4),
                 (36.0, 5, 0.6, 0.7), (44.0, 6, 0.3, 0.9), (50.0, 7, 0.08, 0.5), (58.0, 8, 0.4, 0.6),
                 (72.0, 9, 0.25, 0.4), (78.0, 10, 0.3, 0.8), (88.0, 11, 0.0, 1.0), (92.0, 12, 0.2, 1.0)]
    CHECKPOINTS = [15, 30, 45, 60, 75, 90]

    def __init__(self, tables=None, seed=None):
        self.tables = tables if tables is not None else _load("quadruped_parkour")
        t = self.tables
        self.model = ref.load_model(t)
        self.data = ref.RefData(self.model)
        self.dt = 0.01; self.frame_skip = 10; self.max_episode_steps = 6000
        self.start_pos = np.array([2.0, 0.0, 0.6]); self.finish_pos = np.array([98.0, 0.0, 0.0])
        self.course_width = 20.0
        self.initial_qpos = self.data.qpos.copy(); self.initial_qvel = self.data.qvel.copy()
        self.torso_id = t.name2id("body", "torso")
        self.foot_ids = [t.name2id("body", n) for n in ("fl_foot", "fr_foot", "bl_foot", "br_foot")]
        self.platform_joint_id = t.name2id("joint", "platform_slide")
        self.pendulum_joint_id = t.name2id("joint", "pendulum_swing")
        self.platform_motor_id = t.name2id("actuator", "platform_motor")
        self.pendulum_motor_id = t.name2id("actuator", "pendulum_motor")
        lim = []
        for n in t.names["joint"][1:17]:
            lim.append(80.0 if "hip" in n else 60.0 if "knee" in n else 40.0)
        self.action_high = np.array(lim, np.float32); self.action_low = -self.action_high
        self.np_random = np.random.default_rng(seed)
        self._clear_counters()

    def _clear_counters(self):
        self.step_count = 0; self.episode_reward = 0.0
        self.last_position = self.start_pos.copy(); self.max_forward_progress = 0.0
        self.checkpoints_reached = set(); self.fall_count = 0; self.stuck_counter = 0

    # -- reset: parkour_env.py:314-354; `randomize` lets tests inject the two random draws
    def reset(self, seed=None, randomize=None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        d = self.data
        ref.mj_resetData(self.model, d)
        d.qpos[:] = self.initial_qpos; d.qvel[:] = self.initial_qvel
        d.qpos[0:3] = self.start_pos; d.qpos[3:7] = [1, 0, 0, 0]
        self._clear_counters()
        if randomize is None:
            randomize = (self.np_random.uniform(-1.5, 1.5), self.np_random.uniform(-1.0, 1.0))
        # joint ids used as qpos addresses (parkour_env.py:757-774): lands on bl_knee / bl_ankle
        d.qpos[self.platform_joint_id] = randomize[0]
        d.qpos[self.pendulum_joint_id] = randomize[1]
        ref.mj_step(self.model, d, 10)
        return self._get_observation(), self._get_info()

    def step(self, action):
        d = self.data
        action = np.clip(np.asarray(action, np.float64), self.action_low, self.action_high)
        d.ctrl[:16] = action
        ref.mj_step(self.model, d, self.frame_skip)
        t = self.step_count * self.dt
        d.ctrl[self.platform_motor_id] = 50.0 * np.sin(0.5 * t)
        d.ctrl[self.pendulum_motor_id] = 100.0 * np.sin(0.3 * t)
        obs = self._get_observation()
        reward = self._calculate_reward(action)
        terminated = self._is_terminated()
        truncated = self.step_count >= self.max_episode_steps
        self.step_count += 1
        self.episode_reward += reward
        return obs, reward, terminated, truncated, self._get_info()

    def _foot_contacts(self):
        c = np.zeros(4, np.float32)
        cons = self.data.contact
        for i, fid in enumerate(self.foot_ids):       # body ids compared with geom ids
            for con in cons:
                if con.geom1 == fid or con.geom2 == fid:
                    c[i] = 1.0
                    break
        return c

    def _get_observation(self):
        d = self.data
        obs = np.zeros(95, np.float32)
        obs[0:16] = d.qpos[7:23]; obs[16:32] = d.qvel[6:22]; obs[32:36] = d.qpos[3:7]
        obs[36:39] = d.qvel[0:3]; obs[39:42] = d.qvel[3:6]; obs[42:45] = d.qpos[0:3]
        obs[45:49] = self._foot_contacts()
        body = d.xpos[self.torso_id]
        for i, fid in enumerate(self.foot_ids):
            obs[49 + 3*i:52 + 3*i] = d.xpos[fid] - body
        obs[61:85] = 10.0
        x = body[0]; k = 0
        for (px, typ, hgt, dif) in self.OBSTACLES:
            if px > x:
                obs[85 + 4*k:89 + 4*k] = (px - x, typ, hgt, dif); k += 1
            if k >= 2:
                break
        obs[93] = 0.0; obs[94] = 0.8
        return obs

    def _calculate_reward(self, action):
        d = self.data
        pos = d.xpos[self.torso_id]; x = pos[0]
        reward = -20.0
        progress = x - self.last_position[0]
        if progress > 0:
            reward += progress * 500.0
            self.max_forward_progress = max(self.max_forward_progress, x)
        elif progress < -0.1:
            reward -= 100.0
        for cx in self.CHECKPOINTS:
            if cx not in self.checkpoints_reached and x >= cx:
                self.checkpoints_reached.add(cx); reward += 1000.0
        for (ox, typ, hgt, dif) in self.OBSTACLES:
            key = ("obs", typ)
            if key not in self.checkpoints_reached and x > ox + 2.0:
                self.checkpoints_reached.add(key); reward += 1000.0 + dif * 1000.0
        if x >= self.finish_pos[0]:
            reward += 5000.0
        if abs(d.qpos[3]) > 0.7:
            reward += 100.0
        cc = float(np.sum(self._foot_contacts()))
        if 1 <= cc <= 3:
            reward += 200.0
        reward -= float(np.sum(np.abs(action))) * 0.1
        if pos[2] < 0.2:
            reward -= 2000.0; self.fall_count += 1
        if d.ncon > 8:
            reward -= 500.0
        if abs(progress) < 0.01:
            self.stuck_counter += 1
            if self.stuck_counter > 100:
                reward -= 100.0
        else:
            self.stuck_counter = 0
        self.last_position = pos.copy()
        return reward

    def _is_terminated(self):
        pos = self.data.xpos[self.torso_id]
        return bool(pos[0] >= self.finish_pos[0] or pos[2] < 0.15 or abs(pos[1]) > self.course_width / 2
                    or self.stuck_counter > 1000 or self.fall_count > 3)

    def _get_info(self):
        x = self.data.xpos[self.torso_id][0]
        return dict(step_count=self.step_count, episode_reward=self.episode_reward,
                    max_forward_progress=self.max_forward_progress,
                    checkpoints_reached=len(self.checkpoints_reached), fall_count=self.fall_count,
                    course_completion=min(1.0, max(0.0, (x - self.start_pos[0]) / (self.finish_pos[0] - self.start_pos[0]))))

    # -- state injection used by the parity tests (mirrors b2_set_state / b2_get_state)
    def task_state(self):
        bits = 0
        for k, cx in enumerate(self.CHECKPOINTS):
            bits |= (cx in self.checkpoints_reached) << k
        for k, o in enumerate(self.OBSTACLES):
            bits |= (("obs", o[1]) in self.checkpoints_reached) << (6 + k)
        return dict(step_count=self.step_count, episode_reward=self.episode_reward, last_x=self.last_position[0],
                    max_forward_progress=self.max_forward_progress, checkpoints=bits, fall_count=self.fall_count,
                    stuck_counter=self.stuck_counter)


TASKS = {"quadruped_parkour": QuadrupedParkourRef}
