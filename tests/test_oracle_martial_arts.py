"""The oracle reproduces its committed martial-arts golden vectors and the task restatement keeps the reference's observable
semantics: 113-entry observation, reset moves dummy #1 (qpos[0:7] aliasing, F8), reward terms."""
import os

import numpy as np
import pytest

from oracle import ref
from oracle.tasks_ref import HumanoidMartialArtsRef

GOLD = os.path.join(os.path.dirname(__file__), "golden", "humanoid_martial_arts.npz")


@pytest.fixture(scope="module")
def tables():
    from mujoco_gymnasium_environments_b200.tasks import load_tables
    return load_tables("humanoid_martial_arts")


def test_model_dimensions(tables):
    t = tables
    assert (t.nq, t.nv, t.nu, t.nbody, t.njnt, t.ngeom, t.npair, t.ntree) == (50, 47, 28, 19, 32, 26, 294, 4)
    assert t.integrator == 0 and t.solver == 2 and t.iterations == 50 and abs(t.timestep - 0.01667) < 1e-12
    assert abs(t.tolerance - 1e-10) < 1e-20


def test_physics_golden(tables):
    gold = np.load(GOLD)
    om = ref.load_model(tables)
    for k in (0, 4):
        d = ref.RefData(om)
        d.qpos[:] = gold["qpos"][k]; d.qvel[:] = gold["qvel"][k]; d.ctrl[:] = gold["ctrl"][k]; d.qacc_warmstart[:] = gold["warm"][k]
        ref.mj_forward(om, d)
        assert d.ncon == gold["ncon"][k] and d.nefc == gold["nefc"][k]
        assert [(c.geom1, c.geom2) for c in d.contact] == [tuple(p) for p in gold["pairs"][k][:d.ncon]]
        assert np.allclose(d.qacc, gold["qacc"][k], rtol=0, atol=1e-7 * (1 + np.abs(gold["qacc"][k]).max()))
        d.qacc_warmstart[:] = gold["warm"][k]       # mj_forward left qacc there (MuJoCo 3.x mj_fwdConstraint); step from the stored warm start
        ref.mj_step(om, d)
        assert np.allclose(d.qpos, gold["qpos1"][k], rtol=0, atol=1e-10)


def test_semantics(tables):
    env = HumanoidMartialArtsRef(tables)
    obs, _ = env.reset(draws=(0.25, -0.5))
    assert obs.shape == (113,)
    # the humanoid stays at its model pose; the "torso" write went to dummy #1 (body 1), whose position is obs[97:100]
    assert np.allclose(obs[0:3], [0, 0, 1.4]) and np.allclose(obs[3:7], [1, 0, 0, 0])
    assert np.allclose(obs[97:100], [0.25, -0.5, 1.4]) and np.allclose(obs[100:103], [-2, 0, 0])
    assert np.allclose(obs[103:106], [0, -2, 1]) and np.all(obs[106:113] == 0)
    o, r, te, tr, _ = env.step(np.zeros(28))
    # upright reward 100 * 1.4/1.75, stance bonus 200 dt (angular... the reference's 'angular' slot is cvel[3:]), distance bonus
    dist = np.hypot(0.25, 0.5)
    assert r == pytest.approx(80.0 + 200 * 0.01667 + 50 * (2.0 - dist), rel=1e-3) and not te
    assert o[112] == np.float32(0.0)          # the observation precedes the reward's stance_stability_time update
    o, r, te, tr, _ = env.step(np.zeros(28))
    assert o[112] == np.float32(0.01667)
    # the humanoid's own free joint is qpos[15:22]: dropping it below 0.5 m terminates with a fall (:610-613)
    env.data.qpos[15 + 2] = 0.3
    o, r, te, tr, _ = env.step(np.zeros(28))
    assert te and env.episode_stats["falls"] == 1 and o[2] == np.float32(0.3)    # xpos of the step's forward pass
