"""The oracle reproduces its committed soccer golden vectors, and the task restatement keeps the reference's observable
semantics: index aliasing at reset, persistent goalkeeper force, accumulating wind force."""
import os

import numpy as np
import pytest

from oracle import ref
from oracle.tasks_ref import HumanoidSoccerRef

GOLD = os.path.join(os.path.dirname(__file__), "golden", "humanoid_soccer.npz")


@pytest.fixture(scope="module")
def tables():
    from mujoco_gymnasium_environments_b200.tasks import load_tables
    return load_tables("humanoid_soccer")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_model_dimensions(tables):
    t = tables          # SURVEY App. A cross-task table and A.4 ids
    assert (t.nq, t.nv, t.nu, t.nbody, t.njnt, t.ngeom, t.npair) == (41, 40, 33, 20, 35, 45, 251)
    assert t.integrator == 0 and t.solver == 0 and t.iterations == 50 and abs(t.timestep - 0.02) < 1e-12
    assert (t.name2id("joint", "goalkeeper_y"), t.name2id("joint", "ball_joint"), t.name2id("joint", "abdomen_y")) == (0, 1, 6)
    assert (t.name2id("body", "opponent_goalkeeper"), t.name2id("body", "ball"), t.name2id("body", "torso")) == (3, 4, 6)
    assert (t.name2id("geom", "field"), t.name2id("geom", "ball_geom"), t.name2id("geom", "right_foot"), t.name2id("geom", "left_foot")) == (0, 29, 41, 44)
    assert abs(t.body_mass[4] - 0.43) < 1e-12                       # ball_geom carries an explicit mass
    assert abs(t.body_mass[3] - 5.0 * 8 * np.prod(t.geom_size[28])) < 1e-9      # synthesised default density 5.0 everywhere else


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
    k = 2
    env = HumanoidSoccerRef(tables)
    o, _ = env.reset(draws=[float(x) for x in gold["task_inject"][k]])
    assert np.allclose(o, gold["task_obs0"][k], atol=1e-6)
    for s in range(gold["task_actions"].shape[0]):
        o, r, te, tr, info = env.step(gold["task_actions"][s, k])
        assert np.allclose(o, gold["task_obs"][s, k], atol=1e-6)
        assert r == pytest.approx(gold["task_rew"][s, k], abs=1e-8) and te == gold["task_term"][s, k]
    assert env.data.qfrc_applied[0] == -100.0


def test_semantics_and_quirks(tables):
    env = HumanoidSoccerRef(tables)
    dr = np.zeros(36); dr[0] = -9.0; dr[1] = 3.0; dr[2] = 0.4; dr[32] = 1.25; dr[33] = 2.0; dr[34] = 0.0
    obs, _ = env.reset(draws=dr)
    d = env.data
    assert obs.shape == (80,) and obs.dtype == np.float32
    # F8: the "robot pose" lands on goalkeeper / ball coordinates; the robot itself starts at the origin and falls
    assert abs(d.xpos[env.torso_id][0]) < 0.05 and abs(d.xpos[env.torso_id][1]) < 0.05
    assert abs(d.xpos[env.ball_id][0] - (-7.0)) < 0.05 and abs(d.xpos[env.ball_id][1] - 3.0) < 0.05
    assert abs(d.qpos[0] - 1.25) < 0.2                                   # goalkeeper_y overwrote robot_x
    q = d.qpos[4:8]; assert abs(np.linalg.norm(q) - 1.0) < 1e-9 and abs(abs(q[2]) - 1.0) < 1e-6     # (0,0,sin(a/2),0) normalised
    assert obs[76] == 1.0 and np.all(np.abs(obs) <= 1.0)
    # wind accumulates while the (stale) ball position is above 0.5 m
    d.qpos[3] = 2.0
    ref.mj_forward(env.model, d)
    env.step(np.zeros(33)); env.step(np.zeros(33))
    assert d.xfrc_applied[env.ball_id, 0] == pytest.approx(2 * 2.0 * 1.0 * 0.1) and d.xfrc_applied[env.ball_id, 1] == pytest.approx(0.0, abs=1e-12)
    o, r, te, tr, info = env.step(np.zeros(33))
    assert o[76] == np.float32(1.0 - 3 / 5000)
