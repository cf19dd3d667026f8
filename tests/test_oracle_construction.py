"""The oracle reproduces its committed construction golden vectors and the task restatement keeps the reference's observable
semantics: 135-entry observation, deterministic termination of two of the four tasks."""
import os

import numpy as np
import pytest

from oracle import ref
from oracle.tasks_ref import HumanoidConstructionRef

GOLD = os.path.join(os.path.dirname(__file__), "golden", "humanoid_construction.npz")


@pytest.fixture(scope="module")
def tables():
    from mujoco_gymnasium_environments_b200.tasks import load_tables
    return load_tables("humanoid_construction")


def test_model_dimensions(tables):
    t = tables
    assert (t.nq, t.nv, t.nu, t.nbody, t.njnt, t.ngeom, t.npair, t.ntree) == (110, 99, 33, 39, 44, 55, 1202, 12)
    assert t.integrator == 1 and t.solver == 2 and t.iterations == 100 and abs(t.timestep - 0.002) < 1e-12


def test_physics_golden(tables):
    gold = np.load(GOLD)
    om = ref.load_model(tables)
    for k in (0, 3):
        d = ref.RefData(om)
        d.qpos[:] = gold["qpos"][k]; d.qvel[:] = gold["qvel"][k]; d.ctrl[:] = gold["ctrl"][k]; d.qacc_warmstart[:] = gold["warm"][k]
        ref.mj_forward(om, d)
        assert d.ncon == gold["ncon"][k] and d.nefc == gold["nefc"][k]
        assert np.allclose(d.qacc, gold["qacc"][k], rtol=0, atol=1e-7 * (1 + np.abs(gold["qacc"][k]).max()))
        d.qacc_warmstart[:] = gold["warm"][k]       # mj_forward left qacc there (MuJoCo 3.x mj_fwdConstraint); step from the stored warm start
        ref.mj_step(om, d)
        assert np.allclose(d.qpos, gold["qpos1"][k], rtol=0, atol=1e-10)


def test_semantics(tables):
    env = HumanoidConstructionRef(tables)
    obs, _ = env.reset(draws=(2, 3.0, 0.25, 25.0))
    assert obs.shape == (135,) and obs[92] == 1.0 and obs[100] == np.float32(0.3) and obs[101] == np.float32(0.25) and obs[102] == np.float32(0.5) and obs[110] == 1.0
    assert np.all(obs[60:90] == 0) and np.all(obs[116:] == 0)
    o, r, te, tr, _ = env.step(np.zeros(33))
    assert r == pytest.approx(30.0 + 1.0 + 5.0) and not te and o[94] == np.float32(1 / 300)
    env.current_step = 299
    o, r, te, tr, _ = env.step(np.zeros(33))
    assert te and env.episode_stats["tasks_completed"] == 1
