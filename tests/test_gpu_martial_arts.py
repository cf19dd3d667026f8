"""Parity of the CUDA path for humanoid_martial_arts (Euler at 16.67 ms with implicit joint damping, Newton-50, 47 dofs in 4
kinematic trees, dynamic cylinders) against the fp64 oracle and the committed golden vectors
(tools/make_golden_martial_arts.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "humanoid_martial_arts.npz")


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-12))


@pytest.fixture(scope="module")
def gpu():
    import torch
    from mujoco_gymnasium_environments_b200 import capi
    from mujoco_gymnasium_environments_b200.tasks import TASKS, load_tables
    t = load_tables("humanoid_martial_arts")
    return dict(torch=torch, capi=capi, tables=t, model=capi.DeviceModel(t, 0), spec=TASKS["humanoid_martial_arts"])


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _batch(gpu, gold):
    torch = gpu["torch"]
    n = gold["qpos"].shape[0]
    b = gpu["capi"].Batch(gpu["model"], gpu["spec"].describe(gpu["tables"]), n, 0, 0)
    f = lambda k: torch.tensor(gold[k], dtype=torch.float32)
    b.set_state(f("qpos"), f("qvel"), f("ctrl"), f("warm"), torch.zeros(n))
    return b


def test_forward_contacts_bit_exact_and_newton_optimum(gpu, gold):
    b = _batch(gpu, gold)
    ncon, geom, dist = b.contacts(48)
    dbg = b.debug_forward()
    gpu["torch"].cuda.synchronize()
    for k in range(gold["qpos"].shape[0]):
        n = int(gold["ncon"][k])
        assert int(ncon[k]) == n
        assert np.array_equal(geom[k, :n].cpu().numpy(), gold["pairs"][k][:n])        # bit-exact pair indices, in order
        assert np.allclose(dist[k, :n].cpu().numpy(), gold["dist"][k][:n], atol=5e-6)
        assert int(dbg["nefc"][k]) == int(gold["nefc"][k])
        assert rel(dbg["qacc"][k].cpu(), gold["qacc"][k]) < 1e-3
    s = b.stats().cpu().numpy()
    assert s[4] == 0 and s[5] == 0 and s[6] == 0
    b.close()


def test_single_euler_step_within_1e4_and_drift(gpu, gold):
    b = _batch(gpu, gold)
    b.physics_step(1)
    st = b.get_state()
    for k in range(gold["qpos"].shape[0]):
        assert rel(st["qpos"][k].cpu(), gold["qpos1"][k]) < 1e-4
        assert rel(st["qvel"][k].cpu(), gold["qvel1"][k]) < 1e-4
    b.physics_step(4)
    st = b.get_state()
    drift = [float(np.max(np.abs(st["qpos"][k].cpu().numpy() - gold["qpos5"][k]))) for k in range(gold["qpos"].shape[0])]
    print("5-step |dq| per state:", np.round(drift, 6))
    assert max(drift) < 2e-3, drift
    b.close()


def test_task_reset_and_steps_match_golden(gpu, gold):
    torch = gpu["torch"]
    n = gold["task_inject"].shape[0]
    b = gpu["capi"].Batch(gpu["model"], gpu["spec"].describe(gpu["tables"]), n, 99, 0)
    obs = torch.zeros((n, 113), device="cuda"); rew = torch.zeros(n, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(n, dtype=torch.uint8, device="cuda")
    b.reset(obs, None, torch.tensor(gold["task_inject"], device="cuda"))
    assert np.max(np.abs(obs.cpu().numpy() - gold["task_obs0"])) < 1e-5            # reset ends with mj_forward, no settle steps
    for s in range(gold["task_actions"].shape[0]):
        b.step(torch.tensor(gold["task_actions"][s], device="cuda"), obs, rew, term, trunc)
        o = obs.cpu().numpy(); g = gold["task_obs"][s]
        assert float(np.max(np.abs(o - g) / (1.0 + np.abs(g)))) < 1e-3, s
        assert np.allclose(rew.cpu().numpy(), gold["task_rew"][s], rtol=1e-4, atol=1e-2), s
        assert np.array_equal(term.cpu().numpy().astype(bool), gold["task_term"][s])
    b.close()


def test_fall_terminates_and_class_api(gpu):
    torch = gpu["torch"]
    desc = gpu["spec"].describe(gpu["tables"])
    b = gpu["capi"].Batch(gpu["model"], desc, 2, 1, 0)
    obs = torch.zeros((2, 113), device="cuda"); rew = torch.zeros(2, device="cuda"); fin = torch.zeros((2, 113), device="cuda")
    term = torch.zeros(2, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(2, dtype=torch.uint8, device="cuda")
    b.reset(obs, None, torch.tensor([[0.1, 0.1], [0.2, -0.2]], device="cuda"))
    st = b.get_state()
    q = st["qpos"].clone(); q[1, 15 + 2] = 0.3                     # the humanoid's own free joint is qpos[15:22]
    b.set_state(q, st["qvel"], st["ctrl"], st["qacc_warmstart"], st["time"])
    b.step(torch.zeros((2, 28), device="cuda"), obs, rew, term, trunc, fin)
    assert term.cpu().tolist() == [0, 1]
    assert abs(float(fin[1, 2]) - 0.3) < 1e-6 and abs(float(obs[1, 2]) - 1.4) < 1e-6     # terminal obs kept, env 1 already reset
    ti, tf = b.get_task_state()
    assert int(ti[1, 0]) == 0 and int(ti[0, 0]) == 1
    b.close()
    from mujoco_gymnasium_environments_b200.envs import HumanoidMartialArtsEnv
    e = HumanoidMartialArtsEnv(render_mode=None)
    o, info = e.reset(seed=0)
    assert o.shape == (113,) and set(info) >= {"episode_stats", "combo_chain", "stance_stability", "current_step"}
    o, r, te, tr, info = e.step(np.zeros(28, np.float32))
    assert isinstance(r, float) and info["current_step"] == 1
    e.close()
