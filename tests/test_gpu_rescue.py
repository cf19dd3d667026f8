"""Parity of the CUDA path for bipedal_rescue (RK4, PGS, 3 175 candidate pairs with the pair tables left in global
memory, explicit gripper pairs, 63 dofs in one island) against the fp64 oracle and the committed golden vectors
(tools/make_golden_rescue.py).  Bounds as in test_gpu_parity.py."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "bipedal_rescue.npz")
REL_1STEP = 1e-4


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-12))


@pytest.fixture(scope="module")
def gpu():
    import torch
    from mujoco_gymnasium_environments_b200 import capi
    from mujoco_gymnasium_environments_b200.tasks import TASKS, load_tables
    t = load_tables("bipedal_rescue")
    return dict(torch=torch, capi=capi, tables=t, model=capi.DeviceModel(t, 0), spec=TASKS["bipedal_rescue"])


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _task_batch_from_gold(gpu, gold):
    """The rescue task kernel (cold pair tables, 48 contacts, 124-row arena) driven at the physics level."""
    torch = gpu["torch"]
    n = gold["qpos"].shape[0]
    b = gpu["capi"].Batch(gpu["model"], gpu["spec"].describe(gpu["tables"]), n, 0, 0)
    f = lambda k: torch.tensor(gold[k], dtype=torch.float32)
    b.set_state(f("qpos"), f("qvel"), f("ctrl"), f("warm"), torch.zeros(n))
    return b


def test_forward_contacts_bit_exact_and_solver_state(gpu, gold):
    from oracle import ref
    b = _task_batch_from_gold(gpu, gold)
    ncon, geom, dist = b.contacts(48)
    dbg = b.debug_forward()
    gpu["torch"].cuda.synchronize()
    om = ref.load_model(gpu["tables"])
    for k in range(gold["qpos"].shape[0]):
        n = int(gold["ncon"][k])
        assert int(ncon[k]) == n
        assert np.array_equal(geom[k, :n].cpu().numpy(), gold["pairs"][k][:n])        # bit-exact pair indices, in order
        assert np.allclose(dist[k, :n].cpu().numpy(), gold["dist"][k][:n], atol=5e-6)
        assert int(dbg["nefc"][k]) == int(gold["nefc"][k])
        assert abs(int(dbg["solver_iter"][k]) - int(gold["iters"][k])) <= 1
        d = ref.RefData(om)
        d.qpos[:] = gold["qpos"][k]; d.qvel[:] = gold["qvel"][k]; d.ctrl[:] = gold["ctrl"][k]; d.qacc_warmstart[:] = gold["warm"][k]
        ref.mj_forward(om, d)
        assert rel(dbg["qfrc_smooth"][k].cpu(), d.qfrc_smooth) < 1e-4
        assert rel(dbg["qacc"][k].cpu(), d.qacc) < 2e-3
    s = b.stats().cpu().numpy()
    assert s[4] == 0 and s[5] == 0 and s[6] == 0
    b.close()


def test_single_rk4_step_within_1e4_and_drift(gpu, gold):
    b = _task_batch_from_gold(gpu, gold)
    b.physics_step(1)
    st = b.get_state()
    for k in range(gold["qpos"].shape[0]):
        assert rel(st["qpos"][k].cpu(), gold["qpos1"][k]) < REL_1STEP
        assert rel(st["qvel"][k].cpu(), gold["qvel1"][k]) < 5 * REL_1STEP     # 4 unconverged PGS-50 solves of ~80 rows per step
    b.physics_step(4)
    st = b.get_state()
    drift = [float(np.max(np.abs(st["qpos"][k].cpu().numpy() - gold["qpos5"][k]))) for k in range(gold["qpos"].shape[0])]
    print("5-step |dq| per state:", np.round(drift, 5))
    assert max(drift[:4]) < 5e-3, drift
    assert np.isfinite(drift).all()
    b.close()


def test_task_reset_and_steps_match_golden(gpu, gold):
    torch = gpu["torch"]
    n = gold["task_inject"].shape[0]
    b = gpu["capi"].Batch(gpu["model"], gpu["spec"].describe(gpu["tables"]), n, 99, 0)
    obs = torch.zeros((n, 102), device="cuda"); rew = torch.zeros(n, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(n, dtype=torch.uint8, device="cuda")
    b.reset(obs, None, torch.tensor(gold["task_inject"], device="cuda"))
    assert np.max(np.abs(obs.cpu().numpy() - gold["task_obs0"])) < 5e-5
    f = lambda k: torch.tensor(gold[k], dtype=torch.float32)
    b.set_state(f("task_q0"), f("task_v0"), torch.zeros((n, 26)), f("task_w0"), torch.zeros(n))
    ti, tf = b.get_task_state()
    tf[:, 4:6] = obs[:, 52:54]                                       # prev_robot_pos.xy
    b.set_task_state(ti, tf)
    for s in range(gold["task_actions"].shape[0]):
        b.step(torch.tensor(gold["task_actions"][s], device="cuda"), obs, rew, term, trunc)
        o = obs.cpu().numpy(); g = gold["task_obs"][s]
        tol = 2e-3 if s < 4 else 5e-2
        err = float(np.max(np.abs(o - g) / (1.0 + np.abs(g))))
        assert err < tol, (s, err)
        r = rew.cpu().numpy(); gr = gold["task_rew"][s]
        assert np.array_equal(np.isinf(r), np.isinf(gr))                 # the +inf first-step reward
        fin = np.isfinite(gr)
        assert np.allclose(r[fin], gr[fin], rtol=1e-3, atol=0.05), s
        assert np.array_equal(term.cpu().numpy().astype(bool), gold["task_term"][s])
    assert b.stats().cpu().numpy()[4] == 0
    b.close()


def test_pickup_dropoff_bookkeeping_and_class_api(gpu):
    torch = gpu["torch"]
    desc = gpu["spec"].describe(gpu["tables"])
    b = gpu["capi"].Batch(gpu["model"], desc, 2, 1, 0)
    obs = torch.zeros((2, 102), device="cuda"); rew = torch.zeros(2, device="cuda")
    term = torch.zeros(2, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(2, dtype=torch.uint8, device="cuda")
    b.reset(obs)
    # put victim 1 right under the robot of env 0: it is picked up (distance < 0.8), +1000 on the following step's count
    st = b.get_state(); q = st["qpos"].clone()
    t = gpu["tables"]; qa = t.jnt_qposadr
    rx, ry = float(q[0, qa[t.name2id("joint", "root_x")]]), float(q[0, qa[t.name2id("joint", "root_y")]])
    vpos = t.body_pos[t.name2id("body", "victim1")]
    q[0, qa[t.name2id("joint", "victim1_x")]] = rx - vpos[0] + 0.3; q[0, qa[t.name2id("joint", "victim1_y")]] = ry - vpos[1]
    b.set_state(q, st["qvel"], st["ctrl"], st["qacc_warmstart"], st["time"])
    b.step(torch.zeros((2, 26), device="cuda"), obs, rew, term, trunc)
    ti, tf = b.get_task_state()
    assert int(ti[0, 2]) == 1 and int(ti[0, 3]) == 1 and int(ti[1, 2]) == 0       # carried mask / carrying flag
    assert float(obs[0, 94]) == 1.0 and float(obs[0, 72]) == 1.0                  # len(carried), victim 1 is_carried
    assert torch.isinf(rew).all()                                                   # first step of the episode: +inf
    b.step(torch.zeros((2, 26), device="cuda"), obs, rew, term, trunc)
    assert torch.isfinite(rew).all()
    b.close()
    from mujoco_gymnasium_environments_b200.envs import BipedalRescueEnv
    e = BipedalRescueEnv(render_mode=None)
    o, info = e.reset(seed=0)
    assert o.shape == (102,) and o.dtype == np.float32 and info["victims_remaining"] == 5
    o, r, te, tr, info = e.step(np.zeros(26, np.float32))
    assert r == float("inf") and isinstance(te, bool) and set(info) >= {"episode_stats", "victims_carried", "energy_remaining", "robot_upright"}
    e.close()
