"""The oracle reproduces its committed dancing golden vectors, and the task restatement keeps the reference's
observable semantics: index aliasing, state leaking across reset, beat clock, move transitions."""
import os

import numpy as np
import pytest

from oracle import ref
from oracle.tasks_ref import HumanoidDancingRef

GOLD = os.path.join(os.path.dirname(__file__), "golden", "humanoid_dancing.npz")


@pytest.fixture(scope="module")
def tables():
    from mujoco_gymnasium_environments_b200.tasks import load_tables
    return load_tables("humanoid_dancing")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_model_dimensions(tables):
    t = tables          # SURVEY App. A cross-task table
    assert (t.nq, t.nv, t.nu, t.nbody, t.njnt, t.ngeom, t.npair) == (29, 29, 29, 16, 29, 17, 106)
    assert t.integrator == 1 and t.solver == 0 and t.iterations == 50 and abs(t.timestep - 0.01667) < 1e-12
    assert t.names["joint"] == HumanoidDancingRef.JOINT_NAMES
    assert (t.name2id("body", "torso"), t.name2id("geom", "right_foot"), t.name2id("geom", "left_foot")) == (2, 13, 16)


def test_physics_golden(tables, gold):
    om = ref.load_model(tables)
    for k in range(gold["qpos"].shape[0]):
        d = ref.RefData(om)
        d.qpos[:] = gold["qpos"][k]; d.qvel[:] = gold["qvel"][k]; d.ctrl[:] = gold["ctrl"][k]; d.qacc_warmstart[:] = gold["warm"][k]
        ref.mj_forward(om, d)
        assert d.ncon == gold["ncon"][k] and d.nefc == gold["nefc"][k]
        assert [(c.geom1, c.geom2) for c in d.contact] == [tuple(p) for p in gold["pairs"][k][:d.ncon].tolist()]
        d.qacc_warmstart[:] = gold["warm"][k]       # mj_forward left qacc there (MuJoCo 3.x mj_fwdConstraint); step from the stored warm start
        ref.mj_step(om, d)
        assert np.allclose(d.qpos, gold["qpos1"][k], rtol=0, atol=1e-11)
        assert np.allclose(d.qvel, gold["qvel1"][k], rtol=0, atol=1e-9)


def test_task_golden(tables, gold):
    k = 0
    env = HumanoidDancingRef(tables)
    inj = gold["task_inject"][k]
    o, _ = env.reset(sequence=[(int(inj[2 * i]), float(inj[2 * i + 1])) for i in range(20)])
    assert np.allclose(o, gold["task_obs0"][k], atol=1e-6)
    d = env.data
    d.qpos[:] = gold["task_q0"][k]; d.qvel[:] = gold["task_v0"][k]; d.qacc_warmstart[:] = gold["task_w0"][k]
    env.prev_joint_vel = d.qvel[6:].copy()
    for s in range(gold["task_actions"].shape[0]):
        o, r, te, tr, info = env.step(gold["task_actions"][s, k])
        assert np.allclose(o, gold["task_obs"][s, k], atol=1e-6)
        assert r == pytest.approx(gold["task_rew"][s, k], abs=1e-8) and te == gold["task_term"][s, k]


def test_semantics_and_quirks(tables):
    env = HumanoidDancingRef(tables, seed=3)
    obs, info = env.reset()
    assert obs.shape == (94,) and obs.dtype == np.float32
    # F8: "root height"/"quaternion w" land on abdomen_z / neck_x, so the torso is yawed by ~1.8 rad at reset
    assert abs(env.data.qpos[2] - 1.8) < 0.05 and abs(env.data.qpos[3] - 1.0) < 0.05
    assert np.all(obs[22:29] == 0) and np.all(obs[52:58] == 0) and np.all(obs[73:76] == 0)
    assert obs[76] == 0.0 and obs[77] == 1.0 and obs[88] == np.float32(0.1) and obs[89] == np.float32(0.5)
    assert obs[78:88].sum() == 1.0 and obs[93] == 1.0
    # beat clock: 0.5 s beats at dt = 0.01667 -> the 30th step wraps
    for _ in range(30):
        env.step(np.zeros(29))
    assert env.beat_count == 1 and env.time_since_last_beat == pytest.approx(30 * 0.01667 - 0.5, abs=1e-12)
    # state that leaks across reset (F12): spotlight and fall_start_step
    env.fall_start_step = 5
    spot = env.spotlight_position.copy()
    env.reset()
    assert hasattr(env, "fall_start_step") and np.array_equal(env.spotlight_position, spot)
    # move transitions append the *next* move to the history
    env2 = HumanoidDancingRef(tables)
    env2.reset(sequence=[(k % 10, 0.05) for k in range(20)])
    for _ in range(4):
        env2.step(np.zeros(29))
    assert env2.current_move_idx >= 1 and env2.move_history[0] == HumanoidDancingRef.MOVES[1]
