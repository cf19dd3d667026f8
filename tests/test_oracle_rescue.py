"""The oracle reproduces its committed rescue golden vectors, and the task restatement keeps the reference's observable
semantics: +inf first-step reward, attributes that survive reset, pickup / drop-off rules."""
import os

import numpy as np
import pytest

from oracle import ref
from oracle.tasks_ref import BipedalRescueRef

GOLD = os.path.join(os.path.dirname(__file__), "golden", "bipedal_rescue.npz")


@pytest.fixture(scope="module")
def tables():
    from mujoco_gymnasium_environments_b200.tasks import load_tables
    return load_tables("bipedal_rescue")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_model_dimensions(tables):
    t = tables          # SURVEY App. A cross-task table
    assert (t.nq, t.nv, t.nu, t.nbody, t.njnt, t.ngeom) == (63, 63, 26, 41, 63, 89)
    assert t.npair == 3175                           # the ten explicit gripper pairs replace their filtered twins
    assert t.integrator == 1 and t.solver == 0 and t.iterations == 50 and abs(t.timestep - 0.02) < 1e-12
    assert t.names["joint"][37:63] == BipedalRescueRef.JOINT_NAMES


def test_physics_golden(tables, gold):
    om = ref.load_model(tables)
    for k in range(0, gold["qpos"].shape[0], 3):
        d = ref.RefData(om)
        d.qpos[:] = gold["qpos"][k]; d.qvel[:] = gold["qvel"][k]; d.ctrl[:] = gold["ctrl"][k]; d.qacc_warmstart[:] = gold["warm"][k]
        ref.mj_forward(om, d)
        assert d.ncon == gold["ncon"][k] and d.nefc == gold["nefc"][k]
        assert [(c.geom1, c.geom2) for c in d.contact] == [tuple(p) for p in gold["pairs"][k][:d.ncon].tolist()]
        d.qacc_warmstart[:] = gold["warm"][k]       # mj_forward left qacc there (MuJoCo 3.x mj_fwdConstraint); step from the stored warm start
        ref.mj_step(om, d)
        assert np.allclose(d.qpos, gold["qpos1"][k], rtol=0, atol=1e-10)
        assert np.allclose(d.qvel, gold["qvel1"][k], rtol=0, atol=1e-8)


def test_semantics_and_quirks(tables):
    env = BipedalRescueRef(tables)
    dr = np.zeros(12); dr[0] = 1.0; dr[1] = -2.0
    obs, _ = env.reset(draws=dr)
    assert obs.shape == (102,) and obs[92] == 1.0 and obs[93] == 1.0 and obs[94] == 0 and obs[95] == 0
    o, r, te, tr, info = env.step(np.zeros(26))
    assert r == float("inf")                          # F12: approach term against closest = inf
    o, r, te, tr, info = env.step(np.zeros(26))
    assert np.isfinite(r) and r == pytest.approx(59.0, abs=5.0)
    # attributes created with hasattr survive reset
    env._fall_timer = 7
    env.reset(draws=dr)
    assert env._fall_timer == 7 and hasattr(env, "_prev_rescued_count")
    # pickup needs distance < 0.8; drop-off inside the safe zone rescues every carried victim
    d = env.data
    env.victims_carried = [0, 3]; env.carrying_victims = True
    d.xpos[env.torso_id][:2] = [19.0, 0.5]
    env._check_victim_interactions()
    assert env.victims_rescued == [0, 3] and env.victims_carried == [] and env.episode_stats["victims_rescued"] == 2
