"""Parity of the CUDA path for humanoid_soccer (Euler + implicit damping, PGS, box-box / capsule-box ground contacts,
joint springs, goalkeeper qfrc_applied, wind xfrc_applied) against the fp64 oracle and the committed golden vectors
(tools/make_golden_soccer.py).  Bounds as in test_gpu_parity.py."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "humanoid_soccer.npz")
REL_1STEP = 1e-4
DRIFT_10 = 5e-3


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-12))


@pytest.fixture(scope="module")
def gpu():
    import torch
    from mujoco_gymnasium_environments_b200 import capi
    from mujoco_gymnasium_environments_b200.tasks import TASKS, load_tables
    t = load_tables("humanoid_soccer")
    return dict(torch=torch, capi=capi, tables=t, model=capi.DeviceModel(t, 0), spec=TASKS["humanoid_soccer"])


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _batch_from_gold(gpu, gold):
    torch = gpu["torch"]
    n = gold["qpos"].shape[0]
    b = gpu["capi"].Batch(gpu["model"], None, n, 0, 0)
    f = lambda k: torch.tensor(gold[k], dtype=torch.float32)
    b.set_state(f("qpos"), f("qvel"), f("ctrl"), f("warm"), torch.zeros(n))
    return b


def test_forward_contacts_bit_exact_and_solver_state(gpu, gold):
    from oracle import ref
    b = _batch_from_gold(gpu, gold)
    ncon, geom, dist = b.contacts()
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
        assert rel(dbg["qacc"][k].cpu(), d.qacc) < 1e-3
    s = b.stats().cpu().numpy()
    assert s[4] == 0 and s[5] == 0 and s[6] == 0
    b.close()


def test_single_step_within_1e4_and_drift(gpu, gold):
    b = _batch_from_gold(gpu, gold)
    b.physics_step(1)
    st = b.get_state()
    for k in range(gold["qpos"].shape[0]):
        assert rel(st["qpos"][k].cpu(), gold["qpos1"][k]) < REL_1STEP
        assert rel(st["qvel"][k].cpu(), gold["qvel1"][k]) < REL_1STEP
        assert rel(st["qacc_warmstart"][k].cpu(), gold["warm1"][k]) < 1e-3
    b.physics_step(9)
    st = b.get_state()
    drift = [float(np.max(np.abs(st["qpos"][k].cpu().numpy() - gold["qpos10"][k]))) for k in range(gold["qpos"].shape[0])]
    print("10-step |dq| per state:", np.round(drift, 5))
    assert max(drift[:5]) < DRIFT_10, drift           # the calm half of the fixture
    assert np.isfinite(drift).all()
    b.close()


def test_task_reset_and_steps_match_golden(gpu, gold):
    torch = gpu["torch"]
    n = gold["task_inject"].shape[0]
    b = gpu["capi"].Batch(gpu["model"], gpu["spec"].describe(gpu["tables"]), n, 99, 0)
    obs = torch.zeros((n, 80), device="cuda"); rew = torch.zeros(n, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(n, dtype=torch.uint8, device="cuda")
    b.reset(obs, None, torch.tensor(gold["task_inject"], device="cuda"))
    assert np.max(np.abs(obs.cpu().numpy() - gold["task_obs0"])) < 2e-5
    for s in range(gold["task_actions"].shape[0]):
        b.step(torch.tensor(gold["task_actions"][s], device="cuda"), obs, rew, term, trunc)
        o = obs.cpu().numpy(); g = gold["task_obs"][s]
        tol = 5e-4 if s < 10 else 2e-2      # then the falling robot's unconverged PGS lets fp32/fp64 separate (drift bound)
        assert float(np.max(np.abs(o - g))) < tol, (s, float(np.max(np.abs(o - g))))
        assert np.allclose(rew.cpu().numpy(), gold["task_rew"][s], rtol=5e-4, atol=0.2 if s < 10 else 5.0), s
        assert np.array_equal(term.cpu().numpy().astype(bool), gold["task_term"][s])      # bit-exact termination flags
        assert not trunc.any()
    # the goalkeeper force written into qfrc_applied[0] persists (env 2 has the ball behind x = -10)
    # (read back through a one-env oracle-free check: the third env's goalkeeper was pushed with the clipped -100 N)
    assert gold["task_qapp"][-1][2] == -100.0
    b.close()


def test_wind_and_goalkeeper_forces_match_live_oracle(gpu):
    """Ball lifted to z = 2 m behind x = -10: the wind accumulates in xfrc_applied[ball] and the goalkeeper is pushed."""
    torch = gpu["torch"]
    from oracle.tasks_ref import HumanoidSoccerRef
    rng = np.random.default_rng(5)
    inj = np.zeros((1, 36), np.float32)
    inj[0, 0] = -13.0; inj[0, 1] = 2.0; inj[0, 2] = 0.1; inj[0, 3:32] = rng.uniform(-.1, .1, 29); inj[0, 32] = -1.0
    inj[0, 33] = 1.7; inj[0, 34] = 0.6; inj[0, 35] = 0.1
    b = gpu["capi"].Batch(gpu["model"], gpu["spec"].describe(gpu["tables"]), 1, 3, 0)
    obs = torch.zeros((1, 80), device="cuda"); rew = torch.zeros(1, device="cuda")
    term = torch.zeros(1, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(1, dtype=torch.uint8, device="cuda")
    b.reset(obs, None, torch.tensor(inj, device="cuda"))
    env = HumanoidSoccerRef(gpu["tables"]); env.reset(draws=[float(x) for x in inj[0]])
    st = b.get_state()
    q = st["qpos"].clone(); q[0, 3] = 2.0                           # ball z
    b.set_state(q, st["qvel"], st["ctrl"], st["qacc_warmstart"], st["time"])
    d = env.data
    d.qpos[:] = q[0].cpu().numpy().astype(np.float64); d.qvel[:] = st["qvel"][0].cpu().numpy().astype(np.float64)
    d.qacc_warmstart[:] = st["qacc_warmstart"][0].cpu().numpy().astype(np.float64)
    for s in range(6):
        a = (rng.uniform(-1, 1, (1, 33)) * 3).astype(np.float32)
        b.step(torch.tensor(a, device="cuda"), obs, rew, term, trunc)
        ro, rr, rt, _, _ = env.step(a[0])
        assert float(np.max(np.abs(obs[0].cpu().numpy() - ro))) < 5e-4, s
        assert abs(float(rew[0]) - rr) < 1e-3 * max(1.0, abs(rr))
    ti, tf = b.get_task_state()
    assert abs(float(tf[0, 9]) - d.xfrc_applied[env.ball_id, 0]) < 1e-6 and d.xfrc_applied[env.ball_id, 0] != 0.0
    assert d.qfrc_applied[0] != 0.0
    b.close()


def test_bookkeeping_sharding_and_class_api(gpu):
    torch = gpu["torch"]
    desc = gpu["spec"].describe(gpu["tables"])
    def run(seed, offset, n_envs, steps=3):
        b = gpu["capi"].Batch(gpu["model"], desc, n_envs, seed, offset)
        obs = torch.zeros((n_envs, 80), device="cuda"); rew = torch.zeros(n_envs, device="cuda")
        term = torch.zeros(n_envs, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(n_envs, dtype=torch.uint8, device="cuda")
        b.reset(obs)
        g = torch.Generator(device="cuda"); g.manual_seed(3)
        acts = (torch.rand((steps, 64, 33), device="cuda", generator=g) * 2 - 1) * 5.0
        for s in range(steps):
            b.step(acts[s, offset:offset + n_envs].contiguous(), obs, rew, term, trunc)
        out = obs.clone(); b.close()
        return out
    o1 = run(7, 0, 64); o2 = run(7, 0, 64); o3 = run(7, 32, 32); o4 = run(8, 0, 64)
    assert torch.equal(o1, o2) and torch.equal(o1[32:], o3) and not torch.equal(o1, o4)
    from mujoco_gymnasium_environments_b200.vector_env import B200VectorEnv
    from mujoco_gymnasium_environments_b200.envs import HumanoidSoccerEnv
    env = B200VectorEnv("humanoid_soccer", 128, seed=1)
    obs, _ = env.reset(seed=1)
    assert obs.shape == (128, 80) and env.single_action_space.shape == (33,)
    o, r, te, tr, infos = env.step(env.action_space.sample() * 0.05)
    assert torch.isfinite(o).all() and torch.isfinite(r).all() and float(o.abs().max()) <= 1.0
    env.close()
    e = HumanoidSoccerEnv(render_mode=None)
    o, info = e.reset(seed=0)
    assert o.shape == (80,) and o.dtype == np.float32 and set(info) >= {"episode_stats", "ball_position", "robot_position", "goal_distance"}
    o, r, te, tr, info = e.step(np.zeros(33, np.float32))
    assert isinstance(r, float) and isinstance(te, bool) and set(info) >= {"ball_contact", "robot_upright", "goal_scored"}
    e.close()
