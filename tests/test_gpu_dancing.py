"""Parity of the CUDA path for humanoid_dancing (RK4, PGS, self contacts incl. capsule-box / box-box pairs) against the
fp64 oracle and the committed golden vectors (tools/make_golden_dancing.py).  Bounds as in test_gpu_parity.py."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "humanoid_dancing.npz")
REL_1STEP = 1e-4
DRIFT_10 = 5e-3           # |dq| (rad) after 10 RK4 steps = 40 forward passes with unconverged PGS, fp32 vs fp64


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-12))


@pytest.fixture(scope="module")
def gpu():
    import torch
    from mujoco_gymnasium_environments_b200 import capi
    from mujoco_gymnasium_environments_b200.tasks import TASKS, load_tables
    t = load_tables("humanoid_dancing")
    return dict(torch=torch, capi=capi, tables=t, model=capi.DeviceModel(t, 0), spec=TASKS["humanoid_dancing"])


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
        assert np.allclose(dist[k, :n].cpu().numpy(), gold["dist"][k][:n], atol=2e-6)
        assert int(dbg["nefc"][k]) == int(gold["nefc"][k])
        # PGS early exit compares an fp32 improvement sum with tolerance 1e-8: it may trip one sweep apart from fp64
        assert abs(int(dbg["solver_iter"][k]) - int(gold["iters"][k])) <= 1
        d = ref.RefData(om)
        d.qpos[:] = gold["qpos"][k]; d.qvel[:] = gold["qvel"][k]; d.ctrl[:] = gold["ctrl"][k]; d.qacc_warmstart[:] = gold["warm"][k]
        ref.mj_forward(om, d)
        assert rel(dbg["qfrc_smooth"][k].cpu(), d.qfrc_smooth) < 1e-5
        assert rel(dbg["qacc"][k].cpu(), d.qacc) < 1e-3
    s = b.stats().cpu().numpy()
    assert s[4] == 0 and s[5] == 0 and s[6] == 0
    b.close()


def test_single_rk4_step_within_1e4_and_drift(gpu, gold):
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
    assert max(drift[:6]) < DRIFT_10, drift          # the calm half of the fixture (|ctrl| <= 2)
    assert np.isfinite(drift).all()
    b.close()


def test_task_reset_and_steps_match_golden(gpu, gold):
    torch = gpu["torch"]
    n = gold["task_inject"].shape[0]
    b = gpu["capi"].Batch(gpu["model"], gpu["spec"].describe(gpu["tables"]), n, 99, 0)
    obs = torch.zeros((n, 94), device="cuda"); rew = torch.zeros(n, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(n, dtype=torch.uint8, device="cuda")
    b.reset(obs, None, torch.tensor(gold["task_inject"], device="cuda"))
    assert np.max(np.abs(obs.cpu().numpy() - gold["task_obs0"])) < 1e-5
    # both sides continue from the same fp32 post-reset state (the knees rest exactly on their limit after reset: the
    # sign of a 1e-10 residue decides whether a limit row exists, a degenerate state in north_star's sense)
    f = lambda k: torch.tensor(gold[k], dtype=torch.float32)
    b.set_state(f("task_q0"), f("task_v0"), torch.zeros((n, 29)), f("task_w0"), torch.zeros(n))
    ti, tf = b.get_task_state()
    tf[:, 36:59] = f("task_v0")[:, 6:].cuda()
    b.set_task_state(ti, tf)
    worst = 0.0
    for s in range(gold["task_actions"].shape[0]):
        b.step(torch.tensor(gold["task_actions"][s], device="cuda"), obs, rew, term, trunc)
        o = obs.cpu().numpy(); g = gold["task_obs"][s]
        # first 12 steps: tight; later the unconverged PGS lets fp32/fp64 trajectories separate (drift bound)
        tol = 5e-4 if s < 12 else 3e-2
        err = float(np.max(np.abs(o[:3] - g[:3]))); worst = max(worst, err)
        assert err < tol, (s, err)
        assert np.allclose(rew.cpu().numpy()[:3], gold["task_rew"][s][:3], rtol=2e-4, atol=2e-2), s
        assert np.array_equal(term.cpu().numpy().astype(bool), gold["task_term"][s])      # bit-exact termination flags
        assert not trunc.any()
        if s < 2:      # the vigorous dancer (env 3, 4 kN m torques, chaotic) is compared while its state is still close
            assert float(np.max(np.abs(o[3] - g[3]))) < 5e-3, s
    b.close()


def test_task_bookkeeping_autoreset_and_sharding(gpu):
    torch = gpu["torch"]
    spec = gpu["spec"]; desc = spec.describe(gpu["tables"])
    def run(seed, offset, n_envs, steps=4):
        b = gpu["capi"].Batch(gpu["model"], desc, n_envs, seed, offset)
        obs = torch.zeros((n_envs, 94), device="cuda"); rew = torch.zeros(n_envs, device="cuda")
        term = torch.zeros(n_envs, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(n_envs, dtype=torch.uint8, device="cuda")
        b.reset(obs)
        g = torch.Generator(device="cuda"); g.manual_seed(3)
        acts = (torch.rand((steps, 64, 29), device="cuda", generator=g) * 2 - 1) * 4.0
        for s in range(steps):
            b.step(acts[s, offset:offset + n_envs].contiguous(), obs, rew, term, trunc)
        out = obs.clone(); b.close()
        return out
    o1 = run(7, 0, 64); o2 = run(7, 0, 64); o3 = run(7, 32, 32); o4 = run(8, 0, 64)
    assert torch.equal(o1, o2) and torch.equal(o1[32:], o3) and not torch.equal(o1, o4)
    # truncation at 3600 steps -> same-step auto-reset: counters cleared, spotlight and fall_start survive
    b = gpu["capi"].Batch(gpu["model"], desc, 2, 1, 0)
    obs = torch.zeros((2, 94), device="cuda"); fin = torch.zeros((2, 94), device="cuda"); rew = torch.zeros(2, device="cuda")
    term = torch.zeros(2, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(2, dtype=torch.uint8, device="cuda")
    b.reset(obs)
    ti, tf = b.get_task_state()
    ti[0, 0] = 3599; ti[0, 5] = 17; tf[0, 8] = 0.25
    b.set_task_state(ti, tf)
    b.step(torch.zeros((2, 29), device="cuda"), obs, rew, term, trunc, fin)
    assert trunc.cpu().tolist() == [1, 0] and term.cpu().tolist() == [0, 0]
    ti2, tf2 = b.get_task_state()
    assert int(ti2[0, 0]) == 0 and int(ti2[1, 0]) == 1 and int(ti2[0, 4]) == 2
    assert abs(float(tf2[0, 8]) - 0.9 * 0.25) < 1e-6                  # spotlight chased (0,0,5) once, not re-initialised
    assert float(tf2[0, 0]) == 0.0 and float(tf2[0, 1]) == 0.5        # performance score / crowd excitement reset
    assert b.stats().cpu().numpy()[0] == 1.0
    b.close()


def test_vector_env_and_class_api(gpu):
    torch = gpu["torch"]
    from mujoco_gymnasium_environments_b200.vector_env import B200VectorEnv
    from mujoco_gymnasium_environments_b200.envs import HumanoidDancingEnv
    env = B200VectorEnv("humanoid_dancing", 256, seed=1)
    obs, _ = env.reset(seed=1)
    assert obs.shape == (256, 94) and env.single_action_space.shape == (29,)
    for _ in range(3):
        o, r, te, tr, infos = env.step(env.action_space.sample() * 0.02)
    assert torch.isfinite(o).all() and torch.isfinite(r).all()
    st = env.episode_stats()
    assert st["substeps"] >= 256 * 13 and st["nan_resets"] == 0
    env.close()
    e = HumanoidDancingEnv(render_mode=None)
    o, info = e.reset(seed=0)
    assert o.shape == (94,) and o.dtype == np.float32 and info["combo_multiplier"] == 1.0
    o, r, te, tr, info = e.step(np.zeros(29, np.float32))
    assert isinstance(r, float) and isinstance(te, bool) and set(info) >= {"episode_stats", "beat_phase", "combo_multiplier", "crowd_excitement", "performance_score"}
    e.close()
