"""The oracle reproduces its committed arm golden vectors, runs BASELINE.json configs[0] (1 env, random actions, 1000
control steps = 10 000 mj_steps) without a solver warning, and the task restatement keeps the reference's observable
semantics: overlapping observation writes, substring geom-name matching, the pickup state machine."""
import os

import numpy as np
import pytest

from oracle import ref
from oracle.tasks_ref import RoboticArmAssemblyRef

GOLD = os.path.join(os.path.dirname(__file__), "golden", "robotic_arm_assembly.npz")


@pytest.fixture(scope="module")
def tables():
    from mujoco_gymnasium_environments_b200.tasks import load_tables
    return load_tables("robotic_arm_assembly")


def test_model_dimensions(tables):
    t = tables
    assert (t.nq, t.nv, t.nu, t.nbody, t.njnt, t.ngeom, t.npair, t.ntree) == (72, 63, 9, 28, 18, 53, 785, 10)
    assert t.integrator == 0 and t.solver == 2 and t.iterations == 50 and abs(t.timestep - 0.002) < 1e-12
    assert int((np.asarray(t.pair_condim) == 6).sum()) == 18


def test_physics_golden(tables):
    gold = np.load(GOLD)
    om = ref.load_model(tables)
    for k in (0, 2):
        d = ref.RefData(om)
        d.qpos[:] = gold["qpos"][k]; d.qvel[:] = gold["qvel"][k]; d.ctrl[:] = gold["ctrl"][k]; d.qacc_warmstart[:] = gold["warm"][k]
        ref.mj_forward(om, d)
        assert d.ncon == gold["ncon"][k] and d.nefc == gold["nefc"][k]
        assert [c.dim for c in d.contact] == gold["dims"][k][:d.ncon].tolist()
        assert np.allclose(d.qacc, gold["qacc"][k], rtol=0, atol=1e-6 * (1 + np.abs(gold["qacc"][k]).max()))
        d.qacc_warmstart[:] = gold["warm"][k]       # mj_forward left qacc there (MuJoCo 3.x mj_fwdConstraint); step from the stored warm start
        ref.mj_step(om, d)
        assert np.allclose(d.qpos, gold["qpos1"][k], rtol=0, atol=1e-9)
    assert (gold["dims"][2] == 6).sum() == 8          # the crafted grasp: eight condim-6 contacts = 80 rows


def test_geom_name_matching():
    f = RoboticArmAssemblyRef.geom_component
    assert f("gripper_left_pad") == 100 and f("gripper_left_finger") == -1
    assert f("cpu_socket") == 5 and f("pcb_bin_base") == 0 and f("screw_bin_base") == -1 and f("battery_connector") == 6
    assert f("screw3_shaft") == 3 and f("cover_tab2") == 8 and f("table_top") == -1


def test_semantics_and_pickup_state_machine(tables):
    gold = np.load(GOLD)
    env = RoboticArmAssemblyRef(tables)
    obs, _ = env.reset()
    assert obs.shape == (110,) and np.all(obs[19:23] == [1, 0, 0, 0])
    assert np.all(obs[79:87] == 0) and obs[87] == 0 and obs[88] == -1 and np.all(obs[89:104] == 0.5) and np.all(obs[104:110] == 0)
    assert np.allclose(obs[23:26], [-0.6, 0.3, 0.76], atol=0.01) and np.allclose(obs[72:75], [0.6, -0.3, 0.76], atol=0.01)   # pcb, cable
    d = env.data
    d.qpos[:] = gold["craft_qpos"]; d.qvel[:] = gold["craft_qvel"]; d.qacc_warmstart[:] = gold["craft_warm"]
    o, r, te, tr, _ = env.step(gold["craft_actions"][0])
    assert env.held_component == "cpu" and env.task_phase == "pickup" and o[87] == 1 and o[88] == 5 and o[109] == 1
    assert r == pytest.approx(gold["craft_rew"][0])
    o, r, te, tr, _ = env.step(gold["craft_actions"][1])
    assert env.held_component is None and env.component_status["cpu"] == "dropped" and o[88] == -1 and o[109] == 0


def test_config0_1000_random_steps(tables):
    """BASELINE.json configs[0]: robotic_arm_assembly_env, 1 env, random actions, 1000 steps on the CPU."""
    env = RoboticArmAssemblyRef(tables)
    rng = np.random.default_rng(0)
    env.reset()
    total = 0.0; episodes = 0
    for s in range(1000):
        o, r, te, tr, _ = env.step(rng.uniform(env.action_low, env.action_high))
        assert np.all(np.isfinite(o)) and np.isfinite(r)
        total += r
        if te or tr:
            episodes += 1; env.reset()
    assert env.data.nwarn == 0 and episodes >= 0
