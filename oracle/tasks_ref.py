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


class HumanoidDancingRef:
    """humanoid_dancing_env/dancing_env.py restated: __init__ :36-154, _get_model_indices :680-720, reset :763-831,
    step :833-894, _generate_dance_sequence :896-905, _set_initial_pose :907-922, _update_rhythm :924-937,
    _update_visual_effects :939-955, _check_move_transition :957-973, _update_crowd_excitement :975-1002,
    _update_episode_stats :1004-1026, _get_observation :1028-1120, _calculate_reward :1122-1207,
    _check_termination :1209-1234, helpers :1237-1282.  Index aliasing (qpos[7+i] paired with joint i's range, qpos[2]/[3]
    written as "root height / quaternion w") and the state that leaks across reset (fall_start_step, spotlight) are kept."""

    MOVES = ["basic_step", "spin", "jump", "moonwalk", "robot_wave", "freeze", "hip_hop_bounce", "breakdance_toprock",
             "salsa_basic", "ballet_pirouette"]
    DIFFICULTY = dict(zip(MOVES, [1, 2, 2, 3, 2, 1, 2, 3, 2, 4]))
    JOINT_NAMES = ["abdomen_x", "abdomen_y", "abdomen_z", "neck_x", "neck_y", "right_shoulder1", "right_shoulder2",
                   "right_elbow", "right_wrist_x", "right_wrist_y", "right_wrist_z", "left_shoulder1", "left_shoulder2",
                   "left_elbow", "left_wrist_x", "left_wrist_y", "left_wrist_z", "right_hip_x", "right_hip_y",
                   "right_hip_z", "right_knee", "right_ankle_x", "right_ankle_y", "left_hip_x", "left_hip_y",
                   "left_hip_z", "left_knee", "left_ankle_x", "left_ankle_y"]

    def __init__(self, tables=None, seed=None):
        self.tables = tables if tables is not None else _load("humanoid_dancing")
        t = self.tables
        self.model = ref.load_model(t)
        self.data = ref.RefData(self.model)
        self.dt = 0.01667; self.max_episode_steps = 3600; self.current_step = 0
        self.floor_radius = 10.0
        self.beat_interval = 60.0 / 120
        self.time_since_last_beat = 0.0; self.beat_count = 0; self.current_measure = 0
        self.dance_sequence = []; self.current_move_idx = 0; self.move_start_time = 0.0
        self.robot_height = 1.8
        self.performance_score = 0.0; self.combo_multiplier = 1.0
        self.spotlight_position = np.array([0.0, 0.0, 5.0])
        self.crowd_excitement = 0.5
        self.num_joints = int(t.nu)
        self.joint_indices = [t.name2id("joint", n) for n in self.JOINT_NAMES]
        self.torso_id = t.name2id("body", "torso")
        self.right_foot_id = t.name2id("geom", "right_foot"); self.left_foot_id = t.name2id("geom", "left_foot")
        self.floor_id = t.name2id("geom", "dance_floor"); self.stage_id = t.name2id("geom", "stage")
        self.jnt_range = np.asarray(t.jnt_range)
        self.action_low = np.full(self.num_joints, -200.0); self.action_high = np.full(self.num_joints, 200.0)
        self.episode_stats = dict(total_score=0, longest_combo=0, energy_used=0.0, time_on_beat=0.0, crowd_rating=0.0)
        self.prev_joint_vel = None
        self.move_history = []
        self.np_random = np.random.default_rng(seed)

    # -- reset :763-831; `sequence` lets tests inject the 20 (move index, duration) draws
    def reset(self, seed=None, sequence=None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        d = self.data
        ref.mj_resetData(self.model, d)
        self.current_step = 0; self.beat_count = 0; self.current_measure = 0; self.time_since_last_beat = 0.0
        self.performance_score = 0.0; self.combo_multiplier = 1.0
        self.crowd_excitement = 0.5
        if sequence is None:
            sequence = []
            for _ in range(20):
                move = int(self.np_random.integers(0, 10)); duration = self.np_random.uniform(1.0, 3.0)
                sequence.append((move, duration))
        self.dance_sequence = [dict(move=self.MOVES[int(mv)], duration=float(du)) for mv, du in sequence]
        self.current_move_idx = 0; self.move_start_time = 0.0
        self.episode_stats = dict(total_score=0, longest_combo=0, energy_used=0.0, time_on_beat=0.0, crowd_rating=0.0)
        # _set_initial_pose :907-922
        d.qpos[0] = 0.0; d.qpos[1] = 0.0; d.qpos[2] = self.robot_height
        d.qpos[3:7] = [1.0, 0.0, 0.0, 0.0]
        for i, joint_idx in enumerate(self.joint_indices):
            qpos_idx = 7 + i
            if qpos_idx < len(d.qpos) and joint_idx < len(self.jnt_range):
                d.qpos[qpos_idx] = 0.0
        ref.mj_step(self.model, d, 10)
        obs = self._get_observation()
        self.prev_joint_vel = d.qvel[6:].copy()
        self.move_history = []
        return obs, dict(episode_stats=dict(self.episode_stats), beat_phase=0.0, combo_multiplier=self.combo_multiplier)

    def step(self, action):
        d = self.data
        action = np.clip(np.asarray(action, np.float64), self.action_low, self.action_high)
        d.ctrl[:] = action
        # _update_rhythm :924-937
        self.time_since_last_beat += self.dt
        if self.time_since_last_beat >= self.beat_interval:
            self.time_since_last_beat -= self.beat_interval
            self.beat_count += 1
            if self.beat_count % 4 == 0:
                self.current_measure += 1
        # _update_visual_effects :939-955 (torso position of the previous forward pass)
        robot_pos = d.xpos[self.torso_id].copy()
        target = np.array([robot_pos[0], robot_pos[1], 5.0])
        self.spotlight_position = self.spotlight_position + 0.1 * (target - self.spotlight_position)
        ref.mj_step(self.model, d)
        self.current_step += 1
        obs = self._get_observation()
        reward = self._calculate_reward(action)
        terminated = self._check_termination()
        truncated = self.current_step >= self.max_episode_steps
        self._update_episode_stats()
        self._update_crowd_excitement()
        self._check_move_transition()
        info = dict(episode_stats=dict(self.episode_stats), beat_phase=self.time_since_last_beat / self.beat_interval,
                    combo_multiplier=self.combo_multiplier, crowd_excitement=self.crowd_excitement,
                    performance_score=self.performance_score)
        self.prev_joint_vel = d.qvel[6:].copy()
        return obs, reward, terminated, truncated, info

    def _check_move_transition(self):
        current_time = self.current_step * self.dt
        move_elapsed = current_time - self.move_start_time
        if self.current_move_idx < len(self.dance_sequence):
            if move_elapsed >= self.dance_sequence[self.current_move_idx]["duration"]:
                self.current_move_idx += 1
                self.move_start_time = current_time
                if self.current_move_idx < len(self.dance_sequence):
                    self.move_history.append(self.dance_sequence[self.current_move_idx]["move"])

    def _update_crowd_excitement(self):
        on_beat = 0.1 if (self.time_since_last_beat < 0.1 or self.time_since_last_beat > self.beat_interval - 0.1) else 0.0
        combo_factor = min(self.combo_multiplier / 10.0, 1.0) * 0.2
        if self.current_move_idx < len(self.dance_sequence):
            difficulty_factor = self.DIFFICULTY[self.dance_sequence[self.current_move_idx]["move"]] / 4.0 * 0.1
        else:
            difficulty_factor = 0.0
        change = (on_beat + combo_factor + difficulty_factor) * 0.01
        self.crowd_excitement = float(np.clip(self.crowd_excitement + change, 0.0, 1.0))
        self.crowd_excitement *= 0.999

    def _update_episode_stats(self):
        self.episode_stats["energy_used"] += float(np.sum(np.abs(self.data.ctrl))) * self.dt
        beat_phase = self.time_since_last_beat / self.beat_interval
        if beat_phase < 0.1 or beat_phase > 0.9:
            self.episode_stats["time_on_beat"] += self.dt
        self.episode_stats["longest_combo"] = max(self.episode_stats["longest_combo"], int(self.combo_multiplier))
        self.episode_stats["crowd_rating"] = self.crowd_excitement
        self.episode_stats["total_score"] = self.performance_score

    def _foot_contacts(self):
        c = np.zeros(2)
        ground = (self.floor_id, self.stage_id)
        for con in self.data.contact:
            if (con.geom1 == self.right_foot_id and con.geom2 in ground) or (con.geom2 == self.right_foot_id and con.geom1 in ground):
                c[0] = 1.0
            if (con.geom1 == self.left_foot_id and con.geom2 in ground) or (con.geom2 == self.left_foot_id and con.geom1 in ground):
                c[1] = 1.0
        return c

    def _is_robot_upright(self):
        w, x, y, z = self.data.xquat[self.torso_id]
        return (w*w - x*x - y*y + z*z) > 0.7          # rot_mat[2, 2] of mju_quat2Mat

    def _get_observation(self):
        d = self.data
        obs = []
        for i in range(self.num_joints):
            qpos_idx = 7 + i
            if i < len(self.joint_indices) and qpos_idx < len(d.qpos):
                lo, hi = self.jnt_range[self.joint_indices[i]]
                if lo < hi:
                    obs.append(np.clip(2 * (d.qpos[qpos_idx] - lo) / (hi - lo) - 1, -1.0, 1.0))
                else:
                    obs.append(0.0)
            else:
                obs.append(0.0)
        for i in range(self.num_joints):
            obs.append(np.clip(d.qvel[6 + i] / 10.0, -1.0, 1.0) if i < len(d.qvel) - 6 else 0.0)
        obs.extend(d.xquat[self.torso_id])
        obs.extend(np.clip(d.qvel[:3] / 5.0, -1.0, 1.0))
        obs.extend(np.clip(d.qvel[3:6] / 10.0, -1.0, 1.0))
        obs.extend(np.clip(d.subtree_com[self.torso_id] / 10.0, -1.0, 1.0))
        obs.extend(self._foot_contacts())
        obs.extend(np.zeros(3))
        obs.append(self.time_since_last_beat / self.beat_interval)
        obs.append((self.beat_interval - self.time_since_last_beat) / self.beat_interval)
        enc = np.zeros(len(self.MOVES))
        if self.current_move_idx < len(self.dance_sequence):
            enc[self.MOVES.index(self.dance_sequence[self.current_move_idx]["move"])] = 1.0
        obs.extend(enc)
        obs.append(np.clip(self.combo_multiplier / 10.0, 0.0, 1.0))
        obs.append(self.crowd_excitement)
        obs.extend(np.clip((self.spotlight_position - d.xpos[self.torso_id]) / 10.0, -1.0, 1.0))
        obs.append(1.0 - min(self.episode_stats["energy_used"] / 1000.0, 1.0))
        return np.array(obs, dtype=np.float32)

    def _calculate_reward(self, action):
        d = self.data
        reward = 0.0
        beat_phase = self.time_since_last_beat / self.beat_interval
        if beat_phase < 0.1 or beat_phase > 0.9:
            if np.linalg.norm(d.qvel[6:]) > 1.0:
                reward += 100.0
                self.combo_multiplier = min(self.combo_multiplier + 0.1, 10.0)
            else:
                self.combo_multiplier = max(self.combo_multiplier - 0.05, 1.0)
        if self._is_robot_upright():
            reward += 30.0
            if self.prev_joint_vel is not None and np.linalg.norm(d.qvel[6:] - self.prev_joint_vel) > 0.5:
                reward += 15.0
        if self.prev_joint_vel is not None:
            reward += 20.0 * np.exp(-0.1 * np.linalg.norm(d.qvel[6:] - self.prev_joint_vel))
        if len(self.move_history) > 2 and len(set(self.move_history[-3:])) == 3:
            reward += 50.0
        move_elapsed = self.current_step * self.dt - self.move_start_time
        if self.current_move_idx < len(self.dance_sequence):
            mv = self.dance_sequence[self.current_move_idx]
            if move_elapsed > mv["duration"] * 0.8:
                reward += 200.0 * self.DIFFICULTY[mv["move"]]
        used = 0.0
        for i, joint_idx in enumerate(self.joint_indices):
            qpos_idx = 7 + i
            if qpos_idx < len(d.qpos) and joint_idx < len(self.jnt_range):
                lo, hi = self.jnt_range[joint_idx]
                if lo < hi:
                    used += abs(d.qpos[qpos_idx] - (lo + hi) / 2) / (hi - lo)
        if used > 5.0:
            reward += 10.0
        reward += -0.05 * float(np.sum(np.square(action)))
        if not self._is_robot_upright():
            reward += -500.0
            self.combo_multiplier = 1.0
        if 0.2 < beat_phase < 0.8 and np.linalg.norm(d.qvel[6:]) > 3.0:
            reward += -5.0
        if reward > 0:
            reward *= self.combo_multiplier
        self.performance_score += reward
        return float(reward)

    def _check_termination(self):
        if not self._is_robot_upright():
            if not hasattr(self, "fall_start_step"):
                self.fall_start_step = self.current_step
            elif self.current_step - self.fall_start_step > 120:
                return True
        elif hasattr(self, "fall_start_step"):
            delattr(self, "fall_start_step")
        pos = self.data.xpos[self.torso_id]
        if np.linalg.norm(pos[:2]) > self.floor_radius * 1.5:
            return True
        if pos[2] < 0.0 or pos[2] > 5.0:
            return True
        return False


class HumanoidSoccerRef:
    """humanoid_soccer_env/soccer_env.py restated: __init__ :32-118, _get_model_indices :222-263, reset :347-396,
    step :398-452, _randomize_initial_state :454-496, _update_environmental_factors :498-508, _update_goalkeeper
    :506-524, _apply_environmental_effects :526-537, _get_observation :539-631, _calculate_reward :633-690,
    _check_termination :692-716, _update_episode_stats :718-730, helpers :733-833."""

    JOINT_NAMES = ["abdomen_y", "abdomen_z", "abdomen_x", "neck_x", "neck_y",
                   "right_shoulder1", "right_shoulder2", "right_elbow", "right_wrist_y", "right_wrist_x", "right_wrist_z",
                   "left_shoulder1", "left_shoulder2", "left_elbow", "left_wrist_y", "left_wrist_x", "left_wrist_z",
                   "right_hip_x", "right_hip_z", "right_hip_y", "right_knee", "right_ankle_y", "right_ankle_x",
                   "left_hip_x", "left_hip_z", "left_hip_y", "left_knee", "left_ankle_y", "left_ankle_x"]
    ROBOT_PARTS = ["foot", "shin", "thigh", "torso", "head", "hand", "arm"]

    def __init__(self, tables=None, seed=None):
        self.tables = tables if tables is not None else _load("humanoid_soccer")
        t = self.tables
        self.model = ref.load_model(t)
        self.data = ref.RefData(self.model)
        self.dt = 0.02; self.max_episode_steps = 5000; self.current_step = 0
        self.num_joints = int(t.nu)
        self.joint_indices = [t.name2id("joint", n) for n in self.JOINT_NAMES]
        self.torso_id = t.name2id("body", "torso"); self.ball_id = t.name2id("body", "ball")
        self.goalkeeper_id = t.name2id("body", "opponent_goalkeeper")
        self.ball_geom_id = t.name2id("geom", "ball_geom")
        self.right_foot_id = t.name2id("geom", "right_foot"); self.left_foot_id = t.name2id("geom", "left_foot")
        self.ball_joint = t.name2id("joint", "ball_joint"); self.goalkeeper_joint = t.name2id("joint", "goalkeeper_y")
        self.jnt_range = np.asarray(t.jnt_range); self.jnt_qposadr = np.asarray(t.jnt_qposadr); self.jnt_dofadr = np.asarray(t.jnt_dofadr)
        self.geom_names = t.names["geom"]
        self.action_low = np.full(self.num_joints, -150.0); self.action_high = np.full(self.num_joints, 150.0)
        self.wind_strength = 0.0; self.wind_direction = np.array([0.0, 0.0])
        self.goal_scored = False
        self.episode_stats = dict(goals_scored=0, ball_contacts=0, distance_traveled=0.0, time_upright=0.0, max_ball_speed=0.0)
        self.prev_ball_pos = None; self.prev_robot_pos = None
        self.np_random = np.random.default_rng(seed)

    # -- reset :347-396; `draws` lets tests inject the 36 random draws in the reference's order
    def reset(self, seed=None, draws=None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        d = self.data
        ref.mj_resetData(self.model, d)
        self.current_step = 0; self.goal_scored = False
        self.episode_stats = dict(goals_scored=0, ball_contacts=0, distance_traveled=0.0, time_upright=0.0, max_ball_speed=0.0)
        it = iter(draws) if draws is not None else None
        U = (lambda lo, hi: float(next(it))) if it is not None else (lambda lo, hi: self.np_random.uniform(lo, hi))
        # _randomize_initial_state :454-496
        robot_x = U(-15.0, -5.0); robot_y = U(-10.0, 10.0); robot_z = 1.4
        a0 = self.jnt_qposadr[0]
        d.qpos[a0:a0 + 3] = [robot_x, robot_y, robot_z]
        angle = U(-0.5, 0.5)
        d.qpos[a0 + 3:a0 + 7] = [np.cos(angle / 2), 0, 0, np.sin(angle / 2)]
        bq = self.jnt_qposadr[self.ball_joint]
        d.qpos[bq:bq + 3] = [robot_x + 2.0, robot_y, 0.15]
        for joint_idx in self.joint_indices:
            if joint_idx < len(d.qpos):
                lo, hi = self.jnt_range[joint_idx]
                if lo < hi:
                    noise = U(-0.1, 0.1)
                    d.qpos[self.jnt_qposadr[joint_idx]] = np.clip((lo + hi) / 2 + noise, lo, hi)
        d.qpos[self.jnt_qposadr[self.goalkeeper_joint]] = U(-2.0, 2.0)
        # _update_environmental_factors :498-508
        self.wind_strength = U(0.0, 2.0)
        wind_angle = U(0, 2 * np.pi)
        self.wind_direction = np.array([np.cos(wind_angle), np.sin(wind_angle)])
        self.field_friction_variation = U(0.05, 0.15)
        ref.mj_step(self.model, d, 10)
        obs = self._get_observation()
        self.prev_ball_pos = d.xpos[self.ball_id].copy(); self.prev_robot_pos = d.xpos[self.torso_id].copy()
        return obs, dict(episode_stats=dict(self.episode_stats))

    def step(self, action):
        d = self.data
        action = np.clip(np.asarray(action, np.float64), self.action_low, self.action_high)
        d.ctrl[:] = action
        # _update_goalkeeper :506-524 (ball position of the previous forward pass)
        ball_pos = d.xpos[self.ball_id]
        if ball_pos[0] < -10.0:
            error = np.clip(ball_pos[1], -3.0, 3.0) - d.qpos[self.jnt_qposadr[self.goalkeeper_joint]]
            d.qfrc_applied[self.goalkeeper_joint] = np.clip(50.0 * error, -100.0, 100.0)
        # _apply_environmental_effects :526-537
        if ball_pos[2] > 0.5:
            d.xfrc_applied[self.ball_id, :2] += self.wind_strength * self.wind_direction * 0.1
        ref.mj_step(self.model, d)
        self.current_step += 1
        obs = self._get_observation()
        reward = self._calculate_reward(action)
        terminated = self._check_termination()
        truncated = self.current_step >= self.max_episode_steps
        robot_pos = d.xpos[self.torso_id]
        self.episode_stats["distance_traveled"] += float(np.linalg.norm(robot_pos - self.prev_robot_pos))
        bd = self.jnt_dofadr[self.ball_joint]
        self.episode_stats["max_ball_speed"] = max(self.episode_stats["max_ball_speed"], float(np.linalg.norm(d.qvel[bd:bd + 3])))
        info = dict(episode_stats=dict(self.episode_stats), ball_position=d.xpos[self.ball_id].copy(),
                    robot_position=robot_pos.copy(), goal_scored=self.goal_scored)
        self.prev_ball_pos = d.xpos[self.ball_id].copy(); self.prev_robot_pos = robot_pos.copy()
        return obs, reward, terminated, truncated, info

    def _is_robot_upright(self):
        w, x, y, z = self.data.xquat[self.torso_id]
        return (w*w - x*x - y*y + z*z) > 0.7

    def _foot_contact_forces(self):
        f = np.zeros(4)
        for con in self.data.contact:
            if (con.geom1 == self.right_foot_id and con.geom2 == 0) or (con.geom2 == self.right_foot_id and con.geom1 == 0):
                f[0] = con.dist; f[1] = np.linalg.norm(con.friction[:2])
            if (con.geom1 == self.left_foot_id and con.geom2 == 0) or (con.geom2 == self.left_foot_id and con.geom1 == 0):
                f[2] = con.dist; f[3] = np.linalg.norm(con.friction[:2])
        return f

    def _check_ball_contact(self):
        for con in self.data.contact:
            if con.geom1 == self.ball_geom_id or con.geom2 == self.ball_geom_id:
                other = con.geom2 if con.geom1 == self.ball_geom_id else con.geom1
                name = self.geom_names[other]
                if name and any(part in name for part in self.ROBOT_PARTS):
                    return True
        return False

    def _get_observation(self):
        d = self.data
        obs = []
        n = min(self.num_joints, 25)
        for i in range(n):
            if i < len(self.joint_indices) and self.joint_indices[i] < len(d.qpos):
                j = self.joint_indices[i]
                lo, hi = self.jnt_range[j]
                obs.append(np.clip(2 * (d.qpos[self.jnt_qposadr[j]] - lo) / (hi - lo) - 1, -1.0, 1.0) if lo < hi else 0.0)
            else:
                obs.append(0.0)
        for i in range(n):
            if i < len(self.joint_indices) and self.joint_indices[i] < len(d.qvel):
                obs.append(np.clip(d.qvel[self.jnt_dofadr[self.joint_indices[i]]] / 10.0, -1.0, 1.0))
            else:
                obs.append(0.0)
        obs.extend(d.xquat[self.torso_id])
        obs.extend(np.clip(d.qvel[:3] / 5.0, -1.0, 1.0))
        obs.extend(np.clip(d.qvel[3:6] / 10.0, -1.0, 1.0))
        robot_pos = d.xpos[self.torso_id]; ball_pos = d.xpos[self.ball_id]
        rel = ball_pos - robot_pos
        obs.extend(np.clip(rel / 30.0, -1.0, 1.0))
        bd = self.jnt_dofadr[self.ball_joint]
        obs.extend(np.clip(d.qvel[bd:bd + 3] / 20.0, -1.0, 1.0))
        obs.extend(np.clip((np.array([24.5, 0.0, 1.22]) - robot_pos) / 30.0, -1.0, 1.0))
        obs.extend(np.clip(self._foot_contact_forces() / 1000.0, -1.0, 1.0))
        obs.extend(np.clip(d.subtree_com[self.torso_id] / 30.0, -1.0, 1.0))
        obs.append(1.0 - self.current_step / self.max_episode_steps)
        obs.append(np.clip(np.linalg.norm(rel) / 50.0, 0.0, 1.0))
        obs.extend(np.clip(d.xpos[self.goalkeeper_id][:2] / 15.0, -1.0, 1.0))
        return np.array(obs, dtype=np.float32)

    def _calculate_reward(self, action):
        d = self.data
        reward = 0.0
        ball_pos = d.xpos[self.ball_id].copy(); robot_pos = d.xpos[self.torso_id].copy()
        if ball_pos[0] > 24.0 and abs(ball_pos[1]) < 3.66 and ball_pos[2] < 2.44:
            reward += 10000.0; self.goal_scored = True; self.episode_stats["goals_scored"] += 1
        if self._check_ball_contact():
            reward += 1000.0; self.episode_stats["ball_contacts"] += 1
        cur = np.linalg.norm(ball_pos - robot_pos)
        if self.prev_ball_pos is not None and self.prev_robot_pos is not None:
            prev = np.linalg.norm(self.prev_ball_pos - self.prev_robot_pos)
            if cur < prev and cur > 2.0:
                reward += 500.0 * (prev - cur)
        if self._is_robot_upright():
            reward += 200.0; self.episode_stats["time_upright"] += self.dt
        goal = np.array([24.5, 0.0, 0.0])
        if self.prev_robot_pos is not None:
            pg = np.linalg.norm(self.prev_robot_pos - goal); cg = np.linalg.norm(robot_pos - goal)
            if cg < pg:
                reward += 100.0 * (pg - cg)
        reward += -0.1 * float(np.sum(np.square(action)))
        if not self._is_robot_upright():
            reward += -1000.0
        if self.prev_ball_pos is not None:
            pbg = np.linalg.norm(self.prev_ball_pos - goal); cbg = np.linalg.norm(ball_pos - goal)
            if cbg < pbg:
                reward += 300.0 * (pbg - cbg)
        return float(reward)

    def _check_termination(self):
        if self.goal_scored:
            return True
        if not self._is_robot_upright() and self.current_step > 100:
            return True
        b = self.data.xpos[self.ball_id]
        if abs(b[0]) > 30.0 or abs(b[1]) > 20.0 or b[2] < -1.0 or b[2] > 10.0:
            return True
        r = self.data.xpos[self.torso_id]
        if abs(r[0]) > 30.0 or abs(r[1]) > 20.0 or r[2] < 0.0 or r[2] > 5.0:
            return True
        return False


class BipedalRescueRef:
    """bipedal_rescue_env/rescue_env.py restated: __init__ :35-119, _get_model_indices :279-322, reset :366-414,
    step :416-471, _randomize_initial_state :473-508, _check_victim_interactions :510-543, _get_observation :545-600,
    _calculate_reward :602-668, _check_termination :670-697, _update_episode_stats :699-706, helpers :708-775."""

    JOINT_NAMES = ["neck_pitch", "neck_yaw", "right_shoulder_pitch", "right_shoulder_roll", "right_elbow", "right_wrist",
                   "right_finger1_joint", "right_finger2_joint", "left_shoulder_pitch", "left_shoulder_roll", "left_elbow",
                   "left_wrist", "left_finger1_joint", "left_finger2_joint", "right_hip_roll", "right_hip_pitch",
                   "right_hip_yaw", "right_knee_joint", "right_ankle_pitch", "right_ankle_roll", "left_hip_roll",
                   "left_hip_pitch", "left_hip_yaw", "left_knee_joint", "left_ankle_pitch", "left_ankle_roll"]
    FIRE = [(np.array([-5.0, -3.0, 0.0]), 1.5), (np.array([8.0, 6.0, 0.0]), 1.2)]

    def __init__(self, tables=None, seed=None):
        self.tables = tables if tables is not None else _load("bipedal_rescue")
        t = self.tables
        self.model = ref.load_model(t)
        self.data = ref.RefData(self.model)
        self.dt = 0.02; self.max_episode_steps = 10000; self.current_step = 0
        self.safe_zone_radius = 3.0; self.safe_zone_pos = np.array([20.0, 0.0, 0.0])
        self.carry_capacity = 2; self.energy_limit = 1000.0; self.current_energy = self.energy_limit
        self.num_victims = 5
        self.victims_rescued = []; self.victims_carried = []
        self.torso_id = t.name2id("body", "torso")
        self.victim_ids = [t.name2id("body", f"victim{i}") for i in range(1, 6)]
        self.joint_indices = [t.name2id("joint", n) for n in self.JOINT_NAMES]
        self.jnt_qposadr = np.asarray(t.jnt_qposadr); self.jnt_dofadr = np.asarray(t.jnt_dofadr)
        self.action_low = np.full(26, -100.0); self.action_high = np.full(26, 100.0)
        self.episode_stats = dict(victims_rescued=0, distance_traveled=0.0, energy_used=0.0, time_to_first_rescue=None, falls=0, collisions=0)
        self.prev_robot_pos = None
        self.closest_victim_distance = float("inf")
        self.carrying_victims = False
        self.np_random = np.random.default_rng(seed)

    def reset(self, seed=None, draws=None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        d = self.data
        ref.mj_resetData(self.model, d)
        self.current_step = 0; self.current_energy = self.energy_limit
        self.victims_rescued = []; self.victims_carried = []; self.carrying_victims = False
        self.closest_victim_distance = float("inf")
        self.episode_stats = dict(victims_rescued=0, distance_traveled=0.0, energy_used=0.0, time_to_first_rescue=None, falls=0, collisions=0)
        it = iter(draws) if draws is not None else None
        U = (lambda lo, hi: float(next(it))) if it is not None else (lambda lo, hi: self.np_random.uniform(lo, hi))
        t = self.tables
        d.qpos[self.jnt_qposadr[t.name2id("joint", "root_x")]] = U(-5.0, 5.0)
        d.qpos[self.jnt_qposadr[t.name2id("joint", "root_y")]] = U(-5.0, 5.0)
        d.qpos[self.jnt_qposadr[t.name2id("joint", "root_z")]] = 1.2
        for i in range(1, self.num_victims + 1):
            jx = t.name2id("joint", f"victim{i}_x"); jy = t.name2id("joint", f"victim{i}_y")
            if jx >= 0 and jy >= 0:
                xo = U(-1.0, 1.0); yo = U(-1.0, 1.0)
                d.qpos[self.jnt_qposadr[jx]] += xo; d.qpos[self.jnt_qposadr[jy]] += yo
        ref.mj_step(self.model, d, 10)
        obs = self._get_observation()
        self.prev_robot_pos = d.xpos[self.torso_id].copy()
        return obs, dict(episode_stats=dict(self.episode_stats))

    def step(self, action):
        d = self.data
        action = np.clip(np.asarray(action, np.float64), self.action_low, self.action_high)
        d.ctrl[:len(action)] = action
        energy_cost = float(np.sum(np.abs(action))) * 0.001
        self.current_energy -= energy_cost
        self.episode_stats["energy_used"] += energy_cost
        ref.mj_step(self.model, d)
        self.current_step += 1
        self._check_victim_interactions()
        obs = self._get_observation()
        reward = self._calculate_reward(action)
        terminated = self._check_termination()
        truncated = self.current_step >= self.max_episode_steps
        robot_pos = d.xpos[self.torso_id]
        self.episode_stats["distance_traveled"] += float(np.linalg.norm(robot_pos[:2] - self.prev_robot_pos[:2]))
        info = dict(episode_stats=dict(self.episode_stats), victims_carried=len(self.victims_carried),
                    energy_remaining=self.current_energy, robot_upright=self._is_robot_upright())
        self.prev_robot_pos = robot_pos.copy()
        return obs, reward, terminated, truncated, info

    def _check_victim_interactions(self):
        d = self.data
        robot_pos = d.xpos[self.torso_id]
        if len(self.victims_carried) < self.carry_capacity:
            for i in range(len(self.victim_ids)):
                if i not in self.victims_rescued and i not in self.victims_carried:
                    distance = np.linalg.norm(robot_pos[:2] - d.xpos[self.victim_ids[i]][:2])
                    if distance < 1.0 and distance < 0.8:          # _check_gripper_contact is a second distance test (:753-758)
                        self.victims_carried.append(i); self.carrying_victims = True
        if self.carrying_victims:
            if np.linalg.norm(robot_pos[:2] - self.safe_zone_pos[:2]) < self.safe_zone_radius:
                for v in self.victims_carried:
                    self.victims_rescued.append(v)
                    self.episode_stats["victims_rescued"] += 1
                    if self.episode_stats["time_to_first_rescue"] is None:
                        self.episode_stats["time_to_first_rescue"] = self.current_step * self.dt
                self.victims_carried = []; self.carrying_victims = False

    def _is_robot_upright(self):
        w, x, y, z = self.data.xquat[self.torso_id]
        return (w*w - x*x - y*y + z*z) > 0.7

    def _get_observation(self):
        d = self.data
        obs = []
        for j in self.joint_indices:
            if j < len(d.qpos):
                obs.extend([d.qpos[self.jnt_qposadr[j]], d.qvel[self.jnt_dofadr[j]] if self.jnt_dofadr[j] < len(d.qvel) else 0.0])
            else:
                obs.extend([0.0, 0.0])
        robot_pos = d.xpos[self.torso_id]
        obs.extend(robot_pos); obs.extend(d.xquat[self.torso_id])
        vs = self.jnt_dofadr[self.tables.name2id("joint", "root_x")]
        obs.extend(d.qvel[vs:vs + 6])
        forces = np.zeros(4)
        cons = self.data.contact
        for i in range(min(d.ncon, 10)):
            forces[0] += abs(cons[i].dist)
        obs.extend(forces)
        for i in range(self.num_victims):
            vp = d.xpos[self.victim_ids[i]]
            obs.extend([vp[0], vp[1], 1.0 if i in self.victims_rescued else 0.0, 1.0 if i in self.victims_carried else 0.0])
        obs.extend(self.safe_zone_pos - robot_pos)
        obs.append(self.current_energy / self.energy_limit)
        obs.append(1.0 - (self.current_step / self.max_episode_steps))
        obs.append(len(self.victims_carried)); obs.append(len(self.victims_rescued))
        for pos, _ in self.FIRE:
            obs.extend(pos - robot_pos)
        return np.array(obs, dtype=np.float32)

    def _calculate_reward(self, action):
        d = self.data
        reward = 0.0
        if hasattr(self, "_prev_rescued_count"):
            new = len(self.victims_rescued) - self._prev_rescued_count
            if new > 0:
                reward += 5000.0 * new
        self._prev_rescued_count = len(self.victims_rescued)
        if hasattr(self, "_prev_carried_count"):
            new = len(self.victims_carried) - self._prev_carried_count
            if new > 0:
                reward += 1000.0 * new
        self._prev_carried_count = len(self.victims_carried)
        robot_pos = d.xpos[self.torso_id]
        mind = float("inf")
        for i in range(self.num_victims):
            if i not in self.victims_rescued and i not in self.victims_carried:
                mind = min(mind, float(np.linalg.norm(robot_pos[:2] - d.xpos[self.victim_ids[i]][:2])))
        if mind < self.closest_victim_distance and mind < 10.0:
            reward += 100.0 * (self.closest_victim_distance - mind)
        self.closest_victim_distance = mind
        if self.carrying_victims:
            sd = float(np.linalg.norm(robot_pos[:2] - self.safe_zone_pos[:2]))
            if hasattr(self, "_prev_safe_zone_distance") and sd < self._prev_safe_zone_distance:
                reward += 200.0 * (self._prev_safe_zone_distance - sd)
            self._prev_safe_zone_distance = sd
        if self._is_robot_upright():
            reward += 50.0
        else:
            reward += -500.0; self.episode_stats["falls"] += 1
        if float(np.sum(np.abs(action))) * 0.001 < 0.5:
            reward += 10.0
        for pos, radius in self.FIRE:
            if np.linalg.norm(robot_pos[:2] - pos[:2]) < radius:
                reward += -200.0
        cons = d.contact
        if any(abs(cons[i].dist) > 0.1 for i in range(min(d.ncon, 20))):
            reward += -100.0; self.episode_stats["collisions"] += 1
        reward += -1.0
        return float(reward)

    def _check_termination(self):
        if len(self.victims_rescued) == self.num_victims:
            return True
        if not self._is_robot_upright():
            if not hasattr(self, "_fall_timer"):
                self._fall_timer = 0
            self._fall_timer += 1
            if self._fall_timer > 100:
                return True
        else:
            self._fall_timer = 0
        if self.current_energy <= 0:
            return True
        p = self.data.xpos[self.torso_id]
        if abs(p[0]) > 25 or abs(p[1]) > 25:
            return True
        return False


class HumanoidConstructionRef:
    """humanoid_construction_env/construction_env.py restated: __init__ :32-153, reset :547-584, step :586-623,
    _get_observation :625-659, _calculate_reward :661-700, _update_task_progress :702-719, _check_terminated :721-737."""

    TASK_TYPES = ["stack_blocks", "operate_crane", "transport_material", "build_structure"]

    def __init__(self, tables=None, seed=None):
        self.tables = tables if tables is not None else _load("humanoid_construction")
        t = self.tables
        self.model = ref.load_model(t)
        self.data = ref.RefData(self.model)
        self.max_episode_steps = 3000; self.current_step = 0
        self.max_blocks = 20; self.blocks_placed = 0; self.safety_violations = 0; self.hard_hat_on = True
        self.current_task = None; self.task_progress = 0.0
        self.wind_strength = 0.0; self.rain_intensity = 0.0; self.temperature = 20.0
        self.humanoid_id = t.name2id("body", "humanoid")
        self.action_low = np.full(33, -200.0); self.action_high = np.full(33, 200.0)
        self.episode_stats = dict(tasks_completed=0, total_reward=0.0)
        self.np_random = np.random.default_rng(seed)

    def reset(self, seed=None, draws=None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        ref.mj_resetData(self.model, self.data)
        self.current_step = 0; self.blocks_placed = 0; self.safety_violations = 0
        if draws is None:
            draws = (int(self.np_random.integers(0, 4)), self.np_random.uniform(0, 5), self.np_random.uniform(0, 0.5), self.np_random.uniform(15, 35))
        self.current_task = self.TASK_TYPES[int(draws[0])]
        self.task_progress = 0.0
        self.episode_stats = dict(tasks_completed=0, total_reward=0.0)
        self.wind_strength = float(draws[1]); self.rain_intensity = float(draws[2]); self.temperature = float(draws[3])
        return self._get_observation(), dict(task=self.current_task)

    def step(self, action):
        d = self.data
        action = np.clip(np.asarray(action, np.float64), self.action_low, self.action_high)
        d.ctrl[:] = action
        ref.mj_step(self.model, d)
        self.current_step += 1
        self._update_task_progress()
        reward = self._calculate_reward(action)
        terminated = self._check_terminated()
        truncated = self.current_step >= self.max_episode_steps
        obs = self._get_observation()
        self.episode_stats["total_reward"] += reward
        return obs, reward, terminated, truncated, dict(task=self.current_task, task_progress=self.task_progress)

    def _get_observation(self):
        d = self.data
        obs = []
        obs.extend(d.qpos[:30]); obs.extend(d.qvel[:30]); obs.extend([0.0] * 30)
        one = [0.0] * 4; one[self.TASK_TYPES.index(self.current_task)] = 1.0
        obs.extend(one); obs.append(self.task_progress); obs.extend([0.0] * 5)
        obs.append(self.wind_strength / 10.0); obs.append(self.rain_intensity); obs.append(self.temperature / 50.0)
        obs.extend([0.0] * 7); obs.append(float(self.hard_hat_on)); obs.append(float(self.safety_violations) / 10.0)
        obs.extend([0.0] * 3); obs.append(float(self.blocks_placed) / self.max_blocks); obs.extend([0.0] * 19)
        return np.array(obs, dtype=np.float32)

    def _calculate_reward(self, action):
        reward = 0.0
        if self.current_task == "stack_blocks":
            reward += self.task_progress * 500.0
        elif self.current_task == "operate_crane":
            reward += 200.0 * 0.1
        elif self.current_task == "transport_material":
            reward += 300.0 * 0.1
        else:
            reward += self.task_progress * 100
        if self.hard_hat_on:
            reward += 100.0 * 0.01
        reward -= self.safety_violations * 100
        reward += -0.2 * float(np.sum(np.abs(action)))
        z = self.data.xpos[self.humanoid_id][2]
        reward += 50.0 * 0.1 if z > 1.0 else -2000.0
        return float(reward)

    def _update_task_progress(self):
        if self.current_task == "stack_blocks":
            self.task_progress = min(1.0, self.blocks_placed / 5)
        elif self.current_task == "operate_crane":
            self.task_progress = min(1.0, self.current_step / 500)
        elif self.current_task == "transport_material":
            self.task_progress = min(1.0, self.current_step / 300)
        else:
            self.task_progress = min(1.0, self.blocks_placed / 10)

    def _check_terminated(self):
        if self.data.xpos[self.humanoid_id][2] < 0.5:
            return True
        if self.task_progress >= 1.0:
            self.episode_stats["tasks_completed"] += 1
            return True
        return self.safety_violations > 3



class HumanoidMartialArtsRef:
    """humanoid_martial_arts_env/martial_arts_env.py restated: __init__ :40-132, reset :442-487, step :489-523,
    _get_observation :525-560, _calculate_reward :562-606, _check_termination :608-621.  One Euler step of
    16.67 ms per env.step, Newton-50.  ``qpos[0:7]`` is the free joint of dummy #1, so reset moves that dummy,
    not the humanoid; ``cvel[:3]`` is the angular part and ``cvel[3:]`` the linear part of MuJoCo's com-based velocity, used
    by the reference under the opposite names; the observation has 113 entries (the space declares 85)."""

    def __init__(self, tables=None, seed=None):
        self.tables = tables if tables is not None else _load("humanoid_martial_arts")
        t = self.tables
        self.model = ref.load_model(t)
        self.data = ref.RefData(self.model)
        self.dt = 0.01667; self.max_episode_steps = 6000; self.current_step = 0
        self.robot_height = 1.75; self.balance_reward = 100.0
        self.torso = t.name2id("body", "torso"); self.right_hand = t.name2id("body", "right_hand")
        self.left_hand = t.name2id("body", "left_hand"); self.right_foot = t.name2id("body", "right_ankle")
        self.left_foot = t.name2id("body", "left_ankle")
        self.dummy1 = t.name2id("body", "dummy1"); self.dummy2 = t.name2id("body", "dummy2")
        self.ctrl_hi = np.asarray(t.act_ctrlrange, np.float64).reshape(-1, 2)[:, 1].copy()
        self.action_low = np.full(28, -1.0); self.action_high = np.full(28, 1.0)
        self.combo_len = 0; self.stance_stability_time = 0.0; self.technique_accuracy = 0.0; self.active_dummy_idx = 0
        self.episode_stats = dict(techniques_performed=0, falls=0)
        self.np_random = np.random.default_rng(seed)

    def reset(self, seed=None, draws=None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        m, d = self.model, self.data
        ref.mj_resetData(m, d)
        if draws is None:
            draws = (self.np_random.uniform(-0.5, 0.5), self.np_random.uniform(-0.5, 0.5))
        d.qpos[0:3] = [0.0 + float(draws[0]), 0.0 + float(draws[1]), 1.4]
        d.qpos[3:7] = [1, 0, 0, 0]
        self.current_step = 0; self.combo_len = 0; self.stance_stability_time = 0.0; self.technique_accuracy = 0.0
        self.active_dummy_idx = 0
        self.episode_stats = dict(techniques_performed=0, falls=0)
        ref.mj_forward(m, d)
        return self._get_observation(), dict(current_step=0)

    def step(self, action):
        d = self.data
        action = np.clip(np.asarray(action, np.float64), -1.0, 1.0)
        d.ctrl[:] = action * self.ctrl_hi
        ref.mj_step(self.model, d)
        self.current_step += 1
        obs = self._get_observation()
        reward = self._calculate_reward(action)
        terminated = self._check_termination()
        truncated = self.current_step >= self.max_episode_steps
        return obs, reward, terminated, truncated, dict(current_step=self.current_step, stance_stability=self.stance_stability_time)

    def _get_observation(self):
        d = self.data
        obs = []
        obs.extend(d.xpos[self.torso]); obs.extend(d.xquat[self.torso])
        obs.extend(d.cvel[self.torso][:3]); obs.extend(d.cvel[self.torso][3:])
        obs.extend(d.qpos[7:]); obs.extend(d.qvel[6:])
        obs.extend(d.xpos[self.dummy1]); obs.extend(d.xpos[self.dummy2]); obs.extend([0.0, -2.0, 1.0])
        obs.extend([0.0, 0.0, 0.0, 0.0])
        obs.append(self.technique_accuracy); obs.append(self.combo_len); obs.append(self.stance_stability_time)
        return np.array(obs, dtype=np.float32)

    def _calculate_reward(self, action):
        d = self.data
        reward = 0.0
        reward += self.balance_reward * min(1.0, d.xpos[self.torso][2] / self.robot_height)
        if np.linalg.norm(d.cvel[self.right_hand][:3]) > 2.0 or np.linalg.norm(d.cvel[self.left_hand][:3]) > 2.0:
            reward += 500; self.episode_stats["techniques_performed"] += 1
        if np.linalg.norm(d.cvel[self.right_foot][:3]) > 3.0 or np.linalg.norm(d.cvel[self.left_foot][:3]) > 3.0:
            reward += 800; self.episode_stats["techniques_performed"] += 1
        if np.linalg.norm(d.cvel[self.torso][3:]) < 0.5:
            self.stance_stability_time += self.dt
            reward += 200 * self.dt
        reward -= float(np.sum(np.abs(action))) * 0.01
        dummy = d.xpos[self.dummy1 if self.active_dummy_idx == 0 else self.dummy2]
        distance = float(np.linalg.norm(dummy[:2] - d.xpos[self.torso][:2]))
        if distance < 2.0:
            reward += 50 * (2.0 - distance)
        return float(reward)

    def _check_termination(self):
        p = self.data.xpos[self.torso]
        if p[2] < 0.5:
            self.episode_stats["falls"] += 1
            return True
        return bool(abs(p[0]) > 5.5 or abs(p[1]) > 5.5)


class RoboticArmAssemblyRef:
    """robotic_arm_assembly_env/assembly_env.py restated: __init__ :28-95, reset :162-192, _reset_components :194-218,
    step :220-250, _apply_action :252-265, _update_task_state :267-297, _get_gripper_contacts :299-322, _calculate_reward
    :331-387, _get_max_contact_force :389-397, _check_termination :399-417, _get_observation :419-472.
    Ten Euler sub-steps of 2 ms per env.step, Newton-50/1e-10, condim-6 pad pairs.  Geom names are matched by substring in
    assembly order exactly as the reference does (so ``cpu_socket`` on the PCB counts as the CPU and ``pcb_bin_base`` as the
    PCB).  ``list(set(contacts))[0]`` is hash-order dependent in the reference; here the first component in contact order
    is taken.  The observation writes overlap (nine components in an eight-component layout) and are applied in the
    reference's order."""

    SEQ = ["pcb", "screw1", "screw2", "screw3", "screw4", "cpu", "battery", "cable", "cover"]
    TARGETS = {"pcb": [0, 0, 0.74], "cpu": [0, 0, 0.76], "screw1": [-0.08, -0.06, 0.735], "screw2": [0.08, -0.06, 0.735],
               "screw3": [-0.08, 0.06, 0.735], "screw4": [0.08, 0.06, 0.735], "battery": [0.05, 0, 0.77],
               "cable": [-0.05, 0, 0.77], "cover": [0, 0, 0.79]}
    INITIAL = {"pcb": [-0.6, 0.3, 0.76], "cpu": [-0.6, 0, 0.76], "screw1": [-0.6, -0.3, 0.76], "screw2": [-0.58, -0.3, 0.76],
               "screw3": [-0.62, -0.3, 0.76], "screw4": [-0.6, -0.28, 0.76], "battery": [0.6, 0.3, 0.76],
               "cable": [0.6, -0.3, 0.76], "cover": [0.6, 0, 0.76]}
    PHASES = {"idle": 0, "pickup": 1, "transport": 2, "align": 3, "insert": 4}

    def __init__(self, tables=None, seed=None):
        self.tables = tables if tables is not None else _load("robotic_arm_assembly")
        t = self.tables
        self.model = ref.load_model(t)
        self.data = ref.RefData(self.model)
        self.max_episode_steps = 150000; self.skip_frames = 10
        self.assembly_tolerance = 0.002; self.force_threshold = 50.0; self.gentle_force_threshold = 10.0
        self.action_low = np.array([-2.0] * 7 + [0.0, 0.0]); self.action_high = np.array([2.0] * 7 + [100.0, 50.0])
        self.comp_body = {c: t.name2id("body", c) for c in self.SEQ}
        self.comp_qadr = {c: int(t.jnt_qposadr[int(t.body_jntadr[self.comp_body[c]])]) for c in self.SEQ}
        self.ee_site = t.name2id("site", "ee_site")
        self.geom_names = [t.id2name("geom", g) for g in range(int(t.ngeom))]
        self.reset()

    @classmethod
    def geom_component(cls, name):
        """-1: neither, 100: a gripper pad, else the index of the first component whose name is a substring (:306-320)."""
        if name and "gripper" in name and "pad" in name:
            return 100
        for i, comp in enumerate(cls.SEQ):
            if name and comp in name:
                return i
        return -1

    def reset(self, seed=None, draws=None):
        m, d = self.model, self.data
        ref.mj_resetData(m, d)
        d.qpos[0:7] = [0, -0.5, 0.5, 0, 0.5, 0, 0]
        for c in self.SEQ:
            a = self.comp_qadr[c]
            d.qpos[a:a + 3] = self.INITIAL[c]; d.qpos[a + 3:a + 7] = [1, 0, 0, 0]
        self.step_count = 0
        self.assembly_progress = {c: False for c in self.SEQ}
        self.component_status = {c: "in_bin" for c in self.SEQ}
        self.task_phase = "idle"; self.held_component = None; self.cumulative_reward = 0.0
        ref.mj_step(m, d, 10)
        return self._get_observation(), dict(step_count=0)

    def step(self, action):
        d = self.data
        self.step_count += 1
        a = np.clip(np.asarray(action, np.float64), self.action_low, self.action_high)
        d.ctrl[0:7] = a[0:7]
        d.ctrl[7] = a[7] / 1000.0; d.ctrl[8] = a[7] / 1000.0
        ref.mj_step(self.model, d, self.skip_frames)
        self._update_task_state()
        reward = self._calculate_reward()
        self.cumulative_reward += reward
        terminated = self._check_termination()
        truncated = self.step_count >= self.max_episode_steps
        return self._get_observation(), reward, terminated, truncated, dict(step_count=self.step_count, task_phase=self.task_phase)

    def _gripper_contacts(self):
        out = []
        for c in self.data.contact:
            k1 = self.geom_component(self.geom_names[c.geom1]); k2 = self.geom_component(self.geom_names[c.geom2])
            if k1 == 100:
                if 0 <= k2 < 9 and self.SEQ[k2] not in out: out.append(self.SEQ[k2])
            elif k2 == 100:
                if 0 <= k1 < 9 and self.SEQ[k1] not in out: out.append(self.SEQ[k1])
        return out

    def _update_task_state(self):
        contacts = self._gripper_contacts()
        if contacts:
            if self.held_component is None:
                self.held_component = contacts[0]; self.task_phase = "pickup"
                self.component_status[self.held_component] = "held"
            else:
                self.task_phase = "transport"
        elif self.held_component:
            pos = self.data.xpos[self.comp_body[self.held_component]]
            if np.linalg.norm(pos - np.array(self.TARGETS[self.held_component])) < self.assembly_tolerance:
                self.assembly_progress[self.held_component] = True
                self.component_status[self.held_component] = "assembled"; self.task_phase = "insert"
            else:
                self.component_status[self.held_component] = "dropped"; self.task_phase = "idle"
            self.held_component = None
        else:
            self.task_phase = "idle"

    def _max_contact_force(self):
        mx = 0.0
        for c in self.data.contact:
            mx = max(mx, abs(c.dist) * 1000)
        return mx

    def _calculate_reward(self):
        reward = -10.0
        if self.task_phase == "pickup" and self.held_component:
            reward += 1000
        for comp, done in self.assembly_progress.items():
            if done and self.component_status[comp] == "assembled":
                reward += 2000 if comp in ("pcb", "cpu") else (500 if comp.startswith("screw") else 1000)
        if self.held_component:
            pos = self.data.xpos[self.comp_body[self.held_component]]
            dist = float(np.linalg.norm(pos - np.array(self.TARGETS[self.held_component])))
            if dist < 0.05:
                reward += 300 * (1 - dist / 0.05)
        mf = self._max_contact_force()
        if mf > self.force_threshold:
            reward -= 5000
        elif mf < self.gentle_force_threshold:
            reward += 200
        reward += -float(np.sum(np.abs(self.data.qvel[0:7]))) * 10
        for comp, st in self.component_status.items():
            if st == "dropped":
                reward -= 2000
        if all(self.assembly_progress.values()):
            reward += 10000
        return float(reward)

    def _check_termination(self):
        if all(self.assembly_progress.values()):
            return True
        q = self.data.qpos[0:7]
        lo = np.array([-3.14, -2.36, -2.97, -3.14, -2.09, -3.14, -3.14]); hi = np.array([3.14, 0.78, 2.97, 3.14, 2.09, 3.14, 3.14])
        return bool(np.any(q < lo * 0.95) or np.any(q > hi * 0.95))

    def _get_observation(self):
        d = self.data
        obs = np.zeros(110, dtype=np.float32)
        obs[0:7] = d.qpos[0:7]; obs[7:14] = d.qvel[0:7]
        obs[14] = (d.qpos[7] + d.qpos[8]) / 2.0 * 1000
        obs[15] = self._max_contact_force()
        obs[16:19] = d.site_xpos[self.ee_site]; obs[19:23] = [1, 0, 0, 0]
        idx = 23
        for comp in self.SEQ:
            obs[idx:idx + 3] = d.xpos[self.comp_body[comp]]; obs[idx + 3:idx + 7] = [1, 0, 0, 0]
            idx += 7
        for i, comp in enumerate(self.SEQ):
            obs[79 + i] = float(self.assembly_progress[comp])
        obs[87] = float(self.held_component is not None)
        obs[88] = self.SEQ.index(self.held_component) if self.held_component else -1
        obs[89:114] = 0.5
        obs[104:110] = 0
        obs[108] = sum(self.assembly_progress.values()) / len(self.assembly_progress) * 100
        obs[109] = self.PHASES.get(self.task_phase, 0)
        return obs


TASKS = {"quadruped_parkour": QuadrupedParkourRef, "humanoid_dancing": HumanoidDancingRef, "humanoid_soccer": HumanoidSoccerRef,
         "bipedal_rescue": BipedalRescueRef, "humanoid_construction": HumanoidConstructionRef,
         "humanoid_martial_arts": HumanoidMartialArtsRef,
         "robotic_arm_assembly": RoboticArmAssemblyRef}
