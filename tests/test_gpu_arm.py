"""Parity of the CUDA path for robotic_arm_assembly (10 Euler sub-steps of 2 ms, Newton-50, 63 dofs in 10 kinematic trees,
785 candidate pairs of which 18 are **condim 6** = 10-row pyramids with torsional and rolling friction) against the fp64
oracle and the committed golden vectors (tools/make_golden_arm.py).  Two physics states and the task case are crafted
grasps, because random actions never bring a gripper pad onto a component."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "robotic_arm_assembly.npz")


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-12))


@pytest.fixture(scope="module")
def gpu():
    import torch
    from mujoco_gymnasium_environments_b200 import capi
    from mujoco_gymnasium_environments_b200.tasks import TASKS, load_tables
    t = load_tables("robotic_arm_assembly")
    return dict(torch=torch, capi=capi, tables=t, model=capi.DeviceModel(t, 0), spec=TASKS["robotic_arm_assembly"])


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def excess(a, g, sens, slack=None):
    """largest |a - g| beyond the oracle's own response to an fp32-sized perturbation of the state (oracle/twin.py), relative
    to the largest entry of g"""
    from oracle.twin import SLACK
    a = np.asarray(a, np.float64); g = np.asarray(g, np.float64)
    return float(np.max(np.maximum(np.abs(a - g) - (SLACK if slack is None else slack) * np.asarray(sens), 0.0)) / (np.max(np.abs(g)) + 1e-12))


def _batch(gpu, gold):
    torch = gpu["torch"]
    n = gold["qpos"].shape[0]
    b = gpu["capi"].Batch(gpu["model"], gpu["spec"].describe(gpu["tables"]), n, 0, 0)
    f = lambda k: torch.tensor(gold[k], dtype=torch.float32)
    b.set_state(f("qpos"), f("qvel"), f("ctrl"), f("warm"), torch.zeros(n))
    return b


def test_forward_contacts_bit_exact_condim6_rows_and_newton_optimum(gpu, gold):
    b = _batch(gpu, gold)
    ncon, geom, dist = b.contacts(128)
    dbg = b.debug_forward()
    gpu["torch"].cuda.synchronize()
    for k in range(gold["qpos"].shape[0]):
        n = int(gold["ncon"][k])
        assert int(ncon[k]) == n
        assert np.array_equal(geom[k, :n].cpu().numpy(), gold["pairs"][k][:n])        # bit-exact pair indices, in order
        assert np.allclose(dist[k, :n].cpu().numpy(), gold["dist"][k][:n], atol=5e-6)
        assert int(dbg["nefc"][k]) == int(gold["nefc"][k])                              # 4 rows per condim-3, 10 per condim-6 contact
        # states 0/1 come from the reference scene as authored: the arm is mounted through the table top (17 cm penetration,
        # link velocities > 10 rad/s after reset), where fp32 Newton agrees to 1e-3; the crafted grasps agree to 1e-4.  Entries
        # the oracle itself cannot pin (a screw standing on its flat shaft cap, the base plate under the flat shoulder box: MPR
        # returns a rounding-decided point of the flat patch) are compared up to the oracle's own spread
        e = excess(dbg["qacc"][k].cpu(), gold["qacc"][k], gold["qacc_sens"][k])
        print(f"arm state {k}: qacc beyond the oracle's own spread {e:.2e} (plain relative error {rel(dbg['qacc'][k].cpu(), gold['qacc'][k]):.2e})")
        assert e < (3e-3 if k < 2 else 3e-4)
    s = b.stats().cpu().numpy()
    assert s[4] == 0 and s[5] == 0 and s[6] == 0
    b.close()


def test_single_step_within_1e4_and_drift(gpu, gold):
    b = _batch(gpu, gold)
    b.physics_step(1)
    st = b.get_state()
    for k in range(gold["qpos"].shape[0]):
        assert excess(st["qpos"][k].cpu(), gold["qpos1"][k], gold["qpos1_sens"][k]) < 1e-4
        assert excess(st["qvel"][k].cpu(), gold["qvel1"][k], gold["qvel1_sens"][k]) < 1e-4
    b.physics_step(4)
    st = b.get_state()
    from oracle.twin import SLACK
    drift = [float(np.max(np.maximum(np.abs(st["qpos"][k].cpu().numpy() - gold["qpos5"][k]) - SLACK * gold["qpos5_sens"][k], 0))) for k in range(gold["qpos"].shape[0])]
    print("5-step |dq| per state:", np.round(drift, 6))
    assert max(drift) < 2e-3, drift
    b.close()


def _task_batch(gpu, n):
    torch = gpu["torch"]
    b = gpu["capi"].Batch(gpu["model"], gpu["spec"].describe(gpu["tables"]), n, 99, 0)
    return b, torch.zeros((n, 110), device="cuda"), torch.zeros(n, device="cuda"), torch.zeros(n, dtype=torch.uint8, device="cuda"), torch.zeros(n, dtype=torch.uint8, device="cuda")


def test_task_reset_and_steps_match_golden(gpu, gold):
    torch = gpu["torch"]
    b, obs, rew, term, trunc = _task_batch(gpu, 1)
    b.reset(obs, None, None)
    g0 = gold["task_obs0"]
    from oracle.twin import SLACK
    assert float(np.max(np.maximum(np.abs(obs[0].cpu().numpy() - g0) - SLACK * gold["task_obs0_sens"], 0.0) / (1.0 + np.abs(g0)))) < 1e-3            # after the 10 settle steps
    for s in range(gold["task_actions"].shape[0]):
        b.step(torch.tensor(gold["task_actions"][s][None], device="cuda"), obs, rew, term, trunc)
        o = obs[0].cpu().numpy(); g = gold["task_obs"][s]
        from oracle.twin import SLACK
        err = np.maximum(np.abs(o - g) - SLACK * gold["task_obs_sens"][s], 0.0) / (1.0 + np.abs(g))
        print(f"arm step {s}: worst obs deviation {float(np.max(np.abs(o - g) / (1.0 + np.abs(g)))):.2e}, beyond the oracle's own spread {err.max():.2e} "
              f"at index {int(err.argmax())}, reward {float(rew[0]):.2f} / {gold['task_rew'][s]:.2f}")
        # 10 sub-steps per control step of a scene that starts 17 cm inside the table: stated bound 2e-2 over these 6 steps,
        # on top of what an fp32-sized perturbation of the post-reset state does to the oracle's own rollout
        assert float(err.max()) < 2e-2, s
        assert np.allclose(float(rew[0]), gold["task_rew"][s], rtol=5e-3, atol=2.0 + SLACK * float(gold["task_rew_sens"][s])), s
        assert bool(term[0]) == bool(gold["task_term"][s])
        if bool(term[0]):
            break
    b.close()


def test_pickup_state_machine_and_class_api(gpu, gold):
    torch = gpu["torch"]
    b, obs, rew, term, trunc = _task_batch(gpu, 1)
    b.reset(obs, None, None)
    st = b.get_state()
    f = lambda k: torch.tensor(gold[k][None], dtype=torch.float32)
    b.set_state(f("craft_qpos"), f("craft_qvel"), torch.zeros((1, 9)), f("craft_warm"), st["time"])
    for s in range(gold["craft_actions"].shape[0]):
        b.step(torch.tensor(gold["craft_actions"][s][None], device="cuda"), obs, rew, term, trunc)
        ti, tf = b.get_task_state()
        assert int(ti[0, 3]) == int(gold["craft_held"][s]) and int(ti[0, 4]) == int(gold["craft_phase"][s]), s
        assert float(obs[0, 88]) == float(gold["craft_held"][s])
        from oracle.twin import SLACK
        assert np.allclose(float(rew[0]), gold["craft_rew"][s], rtol=2e-3, atol=1.0 + SLACK * float(gold["craft_rew_sens"][s])), s
        if s == 0:
            g = gold["craft_obs"][0]
            assert float(np.max(np.maximum(np.abs(obs[0].cpu().numpy() - g) - SLACK * gold["craft_obs_sens"][0], 0.0) / (1.0 + np.abs(g)))) < 2e-3
    assert int(ti[0, 5 + 5]) == 3                               # the CPU ends up 'dropped'
    b.close()
    from mujoco_gymnasium_environments_b200.envs import RoboticArmAssemblyEnv
    e = RoboticArmAssemblyEnv(render_mode=None, config={"anything": 1})
    o, info = e.reset(seed=0)
    assert o.shape == (110,) and set(info) >= {"step_count", "assembly_progress", "component_status", "task_phase", "held_component"}
    o, r, te, tr, info = e.step(np.zeros(9, np.float32))
    assert isinstance(r, float) and info["step_count"] == 1 and info["task_phase"] in ("idle", "pickup", "transport", "insert")
    e.close()


def test_physics_only_batch_selects_newton_and_condim6_at_run_time(gpu, gold):
    """b2_physics_step / b2_forward on a batch without a task (the kernel that picks the solver from the model and always
    carries the 10-row pyramid path) must agree with the task kernel's compile-time choices."""
    torch = gpu["torch"]
    n = gold["qpos"].shape[0]
    b = gpu["capi"].Batch(gpu["model"], None, n, 0, 0, con_cap=128, row_cap=512)
    f = lambda k: torch.tensor(gold[k], dtype=torch.float32)
    b.set_state(f("qpos"), f("qvel"), f("ctrl"), f("warm"), torch.zeros(n))
    ncon, geom, dist = b.contacts(128)
    dbg = b.debug_forward()
    torch.cuda.synchronize()
    for k in range(n):
        m = int(gold["ncon"][k])
        assert int(ncon[k]) == m and np.array_equal(geom[k, :m].cpu().numpy(), gold["pairs"][k][:m])
        assert int(dbg["nefc"][k]) == int(gold["nefc"][k])
        assert excess(dbg["qacc"][k].cpu(), gold["qacc"][k], gold["qacc_sens"][k]) < (3e-3 if k < 2 else 3e-4)
    b.physics_step(1)
    st = b.get_state()
    for k in range(n):
        assert excess(st["qpos"][k].cpu(), gold["qpos1"][k], gold["qpos1_sens"][k]) < 1e-4 and excess(st["qvel"][k].cpu(), gold["qvel1"][k], gold["qvel1_sens"][k]) < 1e-4
    b.close()
