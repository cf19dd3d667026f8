"""Host-side task registry: model tables + the name lookups each reference env performs in its constructor.

For every task this resolves the reference's ``mj_name2id`` calls into integer ids once
and fills the ``B2TaskDesc`` the kernels read, so no string ever reaches the device.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Dict, List

import numpy as np

from . import capi
from .mjcf import ModelTables, compile_mjcf
from .spaces import Box

_TABLES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tables")


@dataclass
class TaskSpec:
    name: str
    task_id: int
    obs_dim: int
    act_dim: int
    max_episode_steps: int
    frame_skip: int
    render_fps: int
    bytes_per_env_step: int          # SURVEY 8(d) algorithmic HBM bytes
    describe: Callable[[ModelTables], capi.B2TaskDesc]
    action_space: Callable[[ModelTables], Box]
    observation_space: Callable[[ModelTables], Box]
    info_keys: List[str]
    vector_info: Callable = None     # (torch, ti[N, nti], tf[N, ntf], xpos[N, nbody, 3], tables) -> {reference info key: tensor[N] or dict of them}


def load_tables(task: str, assets_root: str | None = None) -> ModelTables:
    """Committed tables by default; recompile from the reference's assets when ``assets_root`` is given."""
    if assets_root is not None:
        from .compose import COMPOSERS
        return compile_mjcf(COMPOSERS[task](assets_root), name=task)
    return ModelTables.load(os.path.join(_TABLES, task + ".npz"))


# ---------------------------------------------------------------------------------------------- quadruped parkour
def _quad_limits(t: ModelTables) -> np.ndarray:
    # parkour_env.py:236-249: +-80 for names containing "hip", 60 "knee", 40 "ankle"
    lim = []
    for n in t.names["joint"][1:17]:
        lim.append(80.0 if "hip" in n else 60.0 if "knee" in n else 40.0)
    return np.array(lim, np.float32)


def _quad_desc(t: ModelTables) -> capi.B2TaskDesc:
    d = capi.B2TaskDesc()
    d.task = capi.TASK_QUADRUPED_PARKOUR
    ids = [t.name2id("body", "torso")] + [t.name2id("body", n) for n in ("fl_foot", "fr_foot", "bl_foot", "br_foot")]
    ids += [t.name2id("joint", "platform_slide"), t.name2id("joint", "pendulum_swing"),
            t.name2id("actuator", "platform_motor"), t.name2id("actuator", "pendulum_motor")]
    for k, v in enumerate(ids):
        d.ids[k] = v
    lim = _quad_limits(t)
    for k in range(16):
        d.act_lo[k] = -lim[k]; d.act_hi[k] = lim[k]
    return d


def _quad_obs_space(t: ModelTables) -> Box:
    # parkour_env.py:260-287
    lo = np.full(95, -np.inf, np.float32); hi = np.full(95, np.inf, np.float32)
    lo[0:16] = -np.pi; hi[0:16] = np.pi; lo[16:32] = -20.0; hi[16:32] = 20.0; lo[32:36] = -1.0; hi[32:36] = 1.0
    lo[48:52] = 0.0; hi[48:52] = 1.0; lo[64:88] = 0.0; hi[64:88] = 10.0
    return Box(lo, hi, dtype=np.float32)


# ---------------------------------------------------------------------------------------------- humanoid dancing
def _dance_desc(t: ModelTables) -> capi.B2TaskDesc:
    # dancing_env.py:680-720: torso body, foot / floor / stage geoms; joint_indices are 0..28 in joint_names order
    d = capi.B2TaskDesc()
    d.task = capi.TASK_HUMANOID_DANCING
    ids = [t.name2id("body", "torso"), t.name2id("geom", "right_foot"), t.name2id("geom", "left_foot"),
           t.name2id("geom", "dance_floor"), t.name2id("geom", "stage")]
    for k, v in enumerate(ids):
        d.ids[k] = v
    for k in range(29):
        d.act_lo[k] = -200.0; d.act_hi[k] = 200.0
    return d


# ---------------------------------------------------------------------------------------------- humanoid soccer
_SOCCER_PARTS = ("foot", "shin", "thigh", "torso", "head", "hand", "arm")          # soccer_env.py:797-798


_SOCCER_JOINTS = ["abdomen_y", "abdomen_z", "abdomen_x", "neck_x", "neck_y",                # soccer_env.py:226-233
                  "right_shoulder1", "right_shoulder2", "right_elbow", "right_wrist_y", "right_wrist_x", "right_wrist_z",
                  "left_shoulder1", "left_shoulder2", "left_elbow", "left_wrist_y", "left_wrist_x", "left_wrist_z",
                  "right_hip_x", "right_hip_z", "right_hip_y", "right_knee", "right_ankle_y", "right_ankle_x",
                  "left_hip_x", "left_hip_z", "left_hip_y", "left_knee", "left_ankle_y", "left_ankle_x"]


def _soccer_desc(t: ModelTables) -> capi.B2TaskDesc:
    d = capi.B2TaskDesc()
    d.task = capi.TASK_HUMANOID_SOCCER
    mask = 0
    for g, name in enumerate(t.names["geom"]):
        if name and any(p in name for p in _SOCCER_PARTS):
            assert g < 64
            mask |= 1 << g
    torso = t.name2id("body", "torso")
    parent = t.body_parentid
    def below(b):
        while b > 0:
            if b == torso:
                return True
            b = int(parent[b])
        return False
    sub = [b for b in range(int(t.nbody)) if below(b)]
    assert sub == list(range(torso, torso + len(sub))), "torso subtree must be contiguous"
    jids = [t.name2id("joint", n) for n in _SOCCER_JOINTS]
    assert jids == list(range(jids[0], jids[0] + 29)), "observed joints must be contiguous"
    lo32 = mask & 0xffffffff; hi32 = (mask >> 32) & 0xffffffff
    as_i32 = lambda u: u - (1 << 32) if u >= (1 << 31) else u
    ids = [torso, t.name2id("body", "ball"), t.name2id("body", "opponent_goalkeeper"), t.name2id("geom", "ball_geom"),
           t.name2id("geom", "right_foot"), t.name2id("geom", "left_foot"), as_i32(lo32), as_i32(hi32), jids[0],
           t.name2id("joint", "goalkeeper_y"), t.name2id("joint", "ball_joint"), sub[0], len(sub)]
    for k, v in enumerate(ids):
        d.ids[k] = v
    for k in range(33):
        d.act_lo[k] = -150.0; d.act_hi[k] = 150.0
    return d


# ---------------------------------------------------------------------------------------------- bipedal rescue
_RESCUE_JOINTS = ["neck_pitch", "neck_yaw", "right_shoulder_pitch", "right_shoulder_roll", "right_elbow", "right_wrist",     # rescue_env.py:298-308
                  "right_finger1_joint", "right_finger2_joint", "left_shoulder_pitch", "left_shoulder_roll", "left_elbow",
                  "left_wrist", "left_finger1_joint", "left_finger2_joint", "right_hip_roll", "right_hip_pitch",
                  "right_hip_yaw", "right_knee_joint", "right_ankle_pitch", "right_ankle_roll", "left_hip_roll",
                  "left_hip_pitch", "left_hip_yaw", "left_knee_joint", "left_ankle_pitch", "left_ankle_roll"]


def _rescue_desc(t: ModelTables) -> capi.B2TaskDesc:
    d = capi.B2TaskDesc()
    d.task = capi.TASK_BIPEDAL_RESCUE
    vb = [t.name2id("body", f"victim{i}") for i in range(1, 6)]
    assert vb == list(range(vb[0], vb[0] + 5)), "victim bodies must be consecutive"
    jids = [t.name2id("joint", n) for n in _RESCUE_JOINTS]
    assert jids == list(range(jids[0], jids[0] + 26)), "observed joints must be consecutive"
    rx = t.name2id("joint", "root_x")
    assert (t.name2id("joint", "root_y"), t.name2id("joint", "root_z")) == (rx + 1, rx + 2)
    vx = [t.name2id("joint", f"victim{i}_x") for i in range(1, 6)]
    vy = [t.name2id("joint", f"victim{i}_y") for i in range(1, 6)]
    assert vx == [vx[0] + 6 * k for k in range(5)] and vy == [x + 1 for x in vx]
    for k, v in enumerate([t.name2id("body", "torso"), vb[0], rx, jids[0], vx[0]]):
        d.ids[k] = v
    for k in range(26):
        d.act_lo[k] = -100.0; d.act_hi[k] = 100.0
    return d


# ---------------------------------------------------------------------------------------------- humanoid construction
def _construction_desc(t: ModelTables) -> capi.B2TaskDesc:
    d = capi.B2TaskDesc()
    d.task = capi.TASK_HUMANOID_CONSTRUCTION
    d.ids[0] = t.name2id("body", "humanoid")             # construction_env.py:512
    for k in range(33):
        d.act_lo[k] = -200.0; d.act_hi[k] = 200.0
    return d


# ---------------------------------------------------------------------------------------------- humanoid martial arts
def _martial_desc(t: ModelTables) -> capi.B2TaskDesc:
    d = capi.B2TaskDesc()
    d.task = capi.TASK_HUMANOID_MARTIAL_ARTS
    for k, name in enumerate(["torso", "right_hand", "left_hand", "right_ankle", "left_ankle", "dummy1", "dummy2"]):
        d.ids[k] = t.name2id("body", name)                # martial_arts_env.py:383-395
    hi = np.asarray(t.act_ctrlrange, np.float64).reshape(-1, 2)[:, 1]
    for k in range(28):
        d.act_lo[k] = -1.0; d.act_hi[k] = float(hi[k])   # act_hi carries actuator_ctrlrange[:, 1] (:495), the action box is [-1, 1]
    return d


# ---------------------------------------------------------------------------------------------- robotic arm assembly
ARM_SEQUENCE = ["pcb", "screw1", "screw2", "screw3", "screw4", "cpu", "battery", "cable", "cover"]       # assembly_env.py:46-49
ARM_TARGETS = {"pcb": [0, 0, 0.74], "cpu": [0, 0, 0.76], "screw1": [-0.08, -0.06, 0.735], "screw2": [0.08, -0.06, 0.735],
               "screw3": [-0.08, 0.06, 0.735], "screw4": [0.08, 0.06, 0.735], "battery": [0.05, 0, 0.77],
               "cable": [-0.05, 0, 0.77], "cover": [0, 0, 0.79]}                                           # :65-75


def arm_geom_component(name: str) -> int:
    """The reference's per-contact geom-name matching (assembly_env.py:299-322) resolved once per geom: 100 for a gripper
    pad, the index of the first component (assembly order) whose name is a substring, else -1."""
    if name and "gripper" in name and "pad" in name:
        return 100
    for i, comp in enumerate(ARM_SEQUENCE):
        if name and comp in name:
            return i
    return -1


def _arm_desc(t: ModelTables) -> capi.B2TaskDesc:
    d = capi.B2TaskDesc()
    d.task = capi.TASK_ROBOTIC_ARM_ASSEMBLY
    for k, comp in enumerate(ARM_SEQUENCE):
        d.ids[k] = t.name2id("body", comp)
        for a in range(3):
            d.aux_f[3 * k + a] = float(ARM_TARGETS[comp][a])
    site = t.name2id("site", "ee_site")                        # :436
    d.ids[9] = int(t.site_bodyid[site])
    for a in range(3):
        d.aux_f[27 + a] = float(np.asarray(t.site_pos).reshape(-1, 3)[site][a])
    assert int(t.ngeom) <= 64
    for g in range(64):
        d.aux_i[g] = arm_geom_component(t.id2name("geom", g)) if g < int(t.ngeom) else -1
    lo = [-2.0] * 7 + [0.0, 0.0]; hi = [2.0] * 7 + [100.0, 50.0]    # :150-152
    for k in range(9):
        d.act_lo[k] = lo[k]; d.act_hi[k] = hi[k]
    return d


# ---------------------------------------------------------------------------------------------- per-env infos (vector API)
# The numeric entries of each reference env's ``info`` dict as device tensors of shape [N], read from the task state rows
# (layouts: the ``ti`` / ``tf`` comments above each task in csrc/b2_tasks.cuh).  String-valued entries (current_move, task,
# task_phase ...) are given as their integer index.
def _popcount(torch, x):
    x = x.to(torch.int64); n = torch.zeros_like(x)
    for k in range(32):
        n += (x >> k) & 1
    return n


def _quad_info(torch, ti, tf, xpos, t):
    x = xpos[:, t.name2id("body", "torso"), 0]
    return {"step_count": ti[:, 0], "episode_reward": tf[:, 0], "max_forward_progress": tf[:, 2], "checkpoints_reached": _popcount(torch, ti[:, 1]),
            "fall_count": ti[:, 2], "course_completion": ((x - 2.0) / 96.0).clamp(0.0, 1.0)}       # parkour_env.py:797-813


def _dance_info(torch, ti, tf, xpos, t):
    d = tf[:, 2:8].contiguous().view(torch.float64)       # time_since_last_beat, combo_multiplier, move_start_time (fp64 on the device)
    return {"episode_stats": {"total_score": tf[:, 0], "longest_combo": ti[:, 6], "energy_used": tf[:, 11], "time_on_beat": tf[:, 12], "crowd_rating": tf[:, 1]},
            "current_move": ti[:, 3] % 10, "beat_phase": d[:, 0] / 0.5, "combo_multiplier": d[:, 1], "crowd_excitement": tf[:, 1],
            "performance_score": tf[:, 0]}                                                        # dancing_env.py:866-874


def _soccer_info(torch, ti, tf, xpos, t):
    goal = torch.tensor([24.5, 0.0, 0.0], device=tf.device)
    return {"episode_stats": {"goals_scored": ti[:, 2], "ball_contacts": ti[:, 3], "distance_traveled": tf[:, 12], "time_upright": tf[:, 11],
                              "max_ball_speed": tf[:, 13]},
            "ball_position": tf[:, 1:4], "robot_position": tf[:, 4:7], "goal_distance": (tf[:, 4:7] - goal).norm(dim=1),
            "goal_scored": ti[:, 1] != 0}                                                         # soccer_env.py:433-441


def _rescue_info(torch, ti, tf, xpos, t):
    return {"episode_stats": {"victims_rescued": ti[:, 11], "distance_traveled": tf[:, 6], "energy_used": tf[:, 7], "time_to_first_rescue": tf[:, 8],
                              "falls": ti[:, 9], "collisions": ti[:, 10]},
            "robot_position": xpos[:, t.name2id("body", "torso")], "victims_remaining": 5 - ti[:, 11], "victims_carried": _popcount(torch, ti[:, 2]),
            "energy_remaining": tf[:, 1]}                                                         # rescue_env.py:454-461


def _construction_info(torch, ti, tf, xpos, t):
    return {"task": ti[:, 1], "task_progress": tf[:, 1], "blocks_placed": torch.zeros_like(ti[:, 0]), "safety_violations": torch.zeros_like(ti[:, 0]),
            "episode_stats": {"tasks_completed": ti[:, 3], "total_reward": tf[:, 0]},
            "weather": {"wind": tf[:, 2], "rain": tf[:, 3], "temperature": tf[:, 4]}}              # construction_env.py:613-620


def _martial_info(torch, ti, tf, xpos, t):
    return {"episode_stats": {"techniques_performed": ti[:, 1], "falls": ti[:, 3]}, "stance_stability": tf[:, 1], "current_step": ti[:, 0]}   # martial_arts_env.py:513-518


def _arm_info(torch, ti, tf, xpos, t):
    return {"step_count": ti[:, 0], "assembly_progress": ti[:, 1], "task_phase": ti[:, 4], "held_component": ti[:, 3], "cumulative_reward": tf[:, 0],
            "success": (ti[:, 1] & 0x1ff) == 0x1ff}                                               # assembly_env.py:243-252


TASKS: Dict[str, TaskSpec] = {
    "quadruped_parkour": TaskSpec(
        name="quadruped_parkour", task_id=capi.TASK_QUADRUPED_PARKOUR, obs_dim=95, act_dim=16, max_episode_steps=6000,
        frame_skip=10, render_fps=100, bytes_per_env_step=1426, describe=_quad_desc,
        action_space=lambda t: Box(-_quad_limits(t), _quad_limits(t), dtype=np.float32),
        observation_space=_quad_obs_space,
        info_keys=["step_count", "episode_reward", "max_forward_progress", "checkpoints_reached", "fall_count",
                   "course_completion"], vector_info=_quad_info),
    "humanoid_dancing": TaskSpec(
        name="humanoid_dancing", task_id=capi.TASK_HUMANOID_DANCING, obs_dim=94, act_dim=29, max_episode_steps=3600,
        frame_skip=1, render_fps=60, bytes_per_env_step=1954, describe=_dance_desc,
        action_space=lambda t: Box(np.full(29, -200.0, np.float32), np.full(29, 200.0, np.float32), dtype=np.float32),
        observation_space=lambda t: Box(np.full(94, -np.inf, np.float32), np.full(94, np.inf, np.float32), dtype=np.float32),
        info_keys=["episode_stats", "current_move", "beat_phase", "combo_multiplier", "crowd_excitement",
                   "performance_score"], vector_info=_dance_info),
    "humanoid_soccer": TaskSpec(
        name="humanoid_soccer", task_id=capi.TASK_HUMANOID_SOCCER, obs_dim=80, act_dim=33, max_episode_steps=5000,
        frame_skip=1, render_fps=50, bytes_per_env_step=1626, describe=_soccer_desc,
        action_space=lambda t: Box(np.full(33, -150.0, np.float32), np.full(33, 150.0, np.float32), dtype=np.float32),
        observation_space=lambda t: Box(np.full(80, -1.0, np.float32), np.full(80, 1.0, np.float32), dtype=np.float32),
        info_keys=["episode_stats", "ball_position", "robot_position", "goal_distance", "ball_contact", "robot_upright",
                   "goal_scored"], vector_info=_soccer_info),
    "bipedal_rescue": TaskSpec(
        name="bipedal_rescue", task_id=capi.TASK_BIPEDAL_RESCUE, obs_dim=102, act_dim=26, max_episode_steps=10000,
        frame_skip=1, render_fps=50, bytes_per_env_step=2190, describe=_rescue_desc,
        action_space=lambda t: Box(np.full(26, -100.0, np.float32), np.full(26, 100.0, np.float32), dtype=np.float32),
        observation_space=lambda t: Box(np.full(102, -np.inf, np.float32), np.full(102, np.inf, np.float32), dtype=np.float32),
        info_keys=["episode_stats", "robot_position", "victims_remaining", "victims_carried", "energy_remaining",
                   "robot_upright"], vector_info=_rescue_info),
    "humanoid_construction": TaskSpec(
        name="humanoid_construction", task_id=capi.TASK_HUMANOID_CONSTRUCTION, obs_dim=135, act_dim=33, max_episode_steps=3000,
        frame_skip=1, render_fps=50, bytes_per_env_step=3222, describe=_construction_desc,
        action_space=lambda t: Box(np.full(33, -200.0, np.float32), np.full(33, 200.0, np.float32), dtype=np.float32),
        # the reference declares 125 entries but returns 135: the actual length is exposed
        observation_space=lambda t: Box(np.full(135, -np.inf, np.float32), np.full(135, np.inf, np.float32), dtype=np.float32),
        info_keys=["task", "task_progress", "blocks_placed", "safety_violations", "episode_stats", "weather"], vector_info=_construction_info),
    "humanoid_martial_arts": TaskSpec(
        name="humanoid_martial_arts", task_id=capi.TASK_HUMANOID_MARTIAL_ARTS, obs_dim=113, act_dim=28, max_episode_steps=6000,
        frame_skip=1, render_fps=60, bytes_per_env_step=1786, describe=_martial_desc,
        action_space=lambda t: Box(np.full(28, -1.0, np.float32), np.full(28, 1.0, np.float32), dtype=np.float32),
        # the reference declares 85 entries but returns 113: the actual length is exposed
        observation_space=lambda t: Box(np.full(113, -np.inf, np.float32), np.full(113, np.inf, np.float32), dtype=np.float32),
        info_keys=["episode_stats", "combo_chain", "stance_stability", "current_step"], vector_info=_martial_info),
    "robotic_arm_assembly": TaskSpec(
        name="robotic_arm_assembly", task_id=capi.TASK_ROBOTIC_ARM_ASSEMBLY, obs_dim=110, act_dim=9, max_episode_steps=150000,
        frame_skip=10, render_fps=50, bytes_per_env_step=2266, describe=_arm_desc,
        action_space=lambda t: Box(np.array([-2.0] * 7 + [0.0, 0.0], np.float32), np.array([2.0] * 7 + [100.0, 50.0], np.float32), dtype=np.float32),
        observation_space=lambda t: Box(np.full(110, -np.inf, np.float32), np.full(110, np.inf, np.float32), dtype=np.float32),
        info_keys=["step_count", "assembly_progress", "component_status", "task_phase", "held_component", "cumulative_reward", "success"], vector_info=_arm_info),
}
