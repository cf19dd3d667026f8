"""Parity of the CUDA path (through the C-ABI) against the fp64 oracle and the committed golden vectors.

Tolerances come from BASELINE.json's north_star: contact pair indices and termination flags bit-exact on
non-degenerate states; single-step qpos/qvel and obs/reward within 1e-4 relative in fp32; stated drift bound for
rollouts.  Relative errors are measured against the max-abs of the reference vector (the state vector mixes metres,
radians and unit quaternion components).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "quadruped_parkour.npz")
REL_1STEP = 1e-4          # north_star single-step bound
DRIFT_10 = 2e-3           # |dq| after 10 physics steps (contact-rich, unconverged PGS-50, fp32 vs fp64)
DRIFT_100_QPOS = 5e-2     # stated drift bound for the first 100 physics steps on calm (|ctrl| <= 0.5) states


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-12))


@pytest.fixture(scope="module")
def gpu():
    import torch
    from mujoco_gymnasium_environments_b200 import capi
    from mujoco_gymnasium_environments_b200.tasks import TASKS, load_tables
    t = load_tables("quadruped_parkour")
    dm = capi.DeviceModel(t, 0)
    return dict(torch=torch, capi=capi, tables=t, model=dm, spec=TASKS["quadruped_parkour"])


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
        d = ref.RefData(om)
        d.qpos[:] = gold["qpos"][k]; d.qvel[:] = gold["qvel"][k]; d.ctrl[:] = gold["ctrl"][k]; d.qacc_warmstart[:] = gold["warm"][k]
        ref.mj_forward(om, d)
        assert int(dbg["nefc"][k]) == d.nefc and int(dbg["solver_iter"][k]) == d.solver_iter
        assert rel(dbg["qfrc_smooth"][k].cpu(), d.qfrc_smooth) < 1e-5
        assert rel(dbg["qacc_smooth"][k].cpu(), d.qacc_smooth) < 1e-4
        assert rel(dbg["qacc"][k].cpu(), d.qacc) < 1e-3
    s = b.stats().cpu().numpy()
    assert s[4] == 0 and s[5] == 0 and s[6] == 0      # nothing dropped: these are non-degenerate, in-capacity states
    b.close()


def test_single_step_state_within_1e4(gpu, gold):
    b = _batch_from_gold(gpu, gold)
    b.physics_step(1)
    st = b.get_state()
    for k in range(gold["qpos"].shape[0]):
        assert rel(st["qpos"][k].cpu(), gold["qpos1"][k]) < REL_1STEP
        assert rel(st["qvel"][k].cpu(), gold["qvel1"][k]) < REL_1STEP
        assert rel(st["qacc_warmstart"][k].cpu(), gold["warm1"][k]) < 1e-3
    b.close()


def test_rollout_drift_bounds(gpu, gold):
    from oracle import ref
    b = _batch_from_gold(gpu, gold)
    b.physics_step(10)
    st = b.get_state()
    for k in range(gold["qpos"].shape[0]):
        assert np.max(np.abs(st["qpos"][k].cpu().numpy() - gold["qpos10"][k])) < DRIFT_10
    b.physics_step(90)
    st = b.get_state()
    om = ref.load_model(gpu["tables"])
    for k in range(4):                                 # the calm half of the fixture
        d = ref.RefData(om)
        d.qpos[:] = gold["qpos"][k]; d.qvel[:] = gold["qvel"][k]; d.ctrl[:] = gold["ctrl"][k]; d.qacc_warmstart[:] = gold["warm"][k]
        ref.mj_step(om, d, 100)
        assert np.max(np.abs(st["qpos"][k].cpu().numpy() - d.qpos)) < DRIFT_100_QPOS
    b.close()


def test_task_reset_and_steps_match_golden(gpu, gold):
    torch = gpu["torch"]
    n = gold["task_inject"].shape[0]
    b = gpu["capi"].Batch(gpu["model"], gpu["spec"].describe(gpu["tables"]), n, 99, 0)
    obs = torch.zeros((n, 95), device="cuda"); rew = torch.zeros(n, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(n, dtype=torch.uint8, device="cuda")
    inject = torch.zeros((n, 4), device="cuda"); inject[:, :2] = torch.tensor(gold["task_inject"], device="cuda")
    b.reset(obs, None, inject)
    assert rel(obs.cpu(), gold["task_obs0"]) < REL_1STEP
    for s in range(gold["task_actions"].shape[0]):
        b.step(torch.tensor(gold["task_actions"][s], device="cuda"), obs, rew, term, trunc)
        o = obs.cpu().numpy()
        assert np.max(np.abs(o - gold["task_obs"][s]) / (1.0 + np.abs(gold["task_obs"][s]))) < 5e-4
        assert np.allclose(rew.cpu().numpy(), gold["task_rew"][s], rtol=1e-4, atol=1e-2)
        assert np.array_equal(term.cpu().numpy().astype(bool), gold["task_term"][s])      # bit-exact termination flags
        assert not trunc.any()
    b.close()


def test_task_matches_live_oracle_with_full_range_actions(gpu):
    """Single control steps (ten `mj_step`s) from the reset state with FULL-range actions on every actuator (x1.0: the bench's
    action distribution), the whole observation compared.  The robot starts at rest, so ten sub-steps are not yet chaotic."""
    torch = gpu["torch"]
    from oracle.tasks_ref import QuadrupedParkourRef
    n = 6
    b = gpu["capi"].Batch(gpu["model"], gpu["spec"].describe(gpu["tables"]), n, 5, 0)
    obs = torch.zeros((n, 95), device="cuda"); rew = torch.zeros(n, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(n, dtype=torch.uint8, device="cuda")
    rng = np.random.default_rng(11)
    inj = np.zeros((n, 4), np.float32); inj[:, 0] = rng.uniform(-1.5, 1.5, n); inj[:, 1] = rng.uniform(-1, 1, n)
    b.reset(obs, None, torch.tensor(inj, device="cuda"))
    st = {k: v.cpu().numpy() for k, v in b.get_state().items()}
    hi = np.asarray(gpu["spec"].action_space(gpu["tables"]).high)
    a = (rng.uniform(-1, 1, (n, 16)) * hi).astype(np.float32)
    b.step(torch.tensor(a, device="cuda"), obs, rew, term, trunc)
    o = obs.cpu().numpy(); worst = 0.0
    for k in range(n):
        env = QuadrupedParkourRef(gpu["tables"])
        env.reset(randomize=(float(inj[k, 0]), float(inj[k, 1])))
        d = env.data                                    # identical fp32 post-reset state on both sides (north_star's protocol)
        d.qpos[:] = st["qpos"][k]; d.qvel[:] = st["qvel"][k]; d.qacc_warmstart[:] = st["qacc_warmstart"][k]
        ro, rr, rt, _, _ = env.step(a[k])
        # joint velocities reach hundreds of rad/s here: compare relative to each block's scale
        e = max(rel(o[k][0:16], ro[0:16]), rel(o[k][16:32], ro[16:32]), rel(o[k][32:45], ro[32:45]), rel(o[k][45:95], ro[45:95]))
        worst = max(worst, e)
        assert e < REL_1STEP, (k, e)                    # north_star's single-step bound, whole observation (measured 2e-6)
        assert float(np.max(np.abs(o[k] - ro) / (1.0 + np.abs(ro)))) < REL_1STEP
        assert bool(term[k]) == rt
        assert abs(float(rew[k]) - rr) < 1e-3 * max(1.0, abs(rr))
    print(f"full-range single control steps: worst block-relative obs error {worst:.2e}")
    s = b.stats().cpu().numpy()
    assert s[4] == 0 and s[5] == 0 and s[6] == 0
    b.close()


def test_full_range_rollout_step_by_step_with_contact_pairs(gpu):
    """25 control steps = 250 `mj_step`s under FULL-range actions (x1.0 on every actuator, the bench's distribution: +-6400 N m on
    the hips, joint speeds of 1e3 rad/s and more).  Such trajectories are chaotic -- a 1e-6 difference grows to O(1) within one
    control step of ten sub-steps -- so the protocol is north_star's single-step one: six envs are driven by the task kernel,
    and at every physics sub-step of every control step the state the GPU is in (fp32 qpos, qvel, ctrl, qacc_warmstart) is given
    to the oracle; both take one mj_step.  Contact count, contact pairs (bit-exact, in order) and row count must agree, qacc and
    the stepped velocity within the bounds below.  The robots go over, pile up 64-80 contacts (wide tier) and blow up to
    |qvel| ~ 1e6 before mj_checkAcc resets them: all of it is compared, nothing may be dropped."""
    torch = gpu["torch"]
    from oracle import ref
    n, T = 6, 25
    b = gpu["capi"].Batch(gpu["model"], gpu["spec"].describe(gpu["tables"]), n, 5, 0)
    b1 = gpu["capi"].Batch(gpu["model"], gpu["spec"].describe(gpu["tables"]), n, 5, 0)      # sub-step probe of the same states
    obs = torch.zeros((n, 95), device="cuda"); rew = torch.zeros(n, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(n, dtype=torch.uint8, device="cuda")
    rng = np.random.default_rng(23)
    inj = np.zeros((n, 4), np.float32); inj[:, 0] = rng.uniform(-1.5, 1.5, n); inj[:, 1] = rng.uniform(-1, 1, n)
    b.reset(obs, None, torch.tensor(inj, device="cuda"))
    hi = np.asarray(gpu["spec"].action_space(gpu["tables"]).high)
    om = ref.load_model(gpu["tables"])
    compared = 0; worst_a = worst_v = 0.0; maxcon = 0; wild = 0
    for s in range(T):
        st = b.get_state()
        a = (rng.uniform(-1, 1, (n, 16)) * hi).astype(np.float32)
        b.step(torch.tensor(a, device="cuda"), obs, rew, term, trunc)
        ctrl = b.get_state()["ctrl"]                                 # what apply_action wrote for this control step
        done = (term | trunc).cpu().numpy().astype(bool)
        b1.set_state(st["qpos"], st["qvel"], ctrl, st["qacc_warmstart"], torch.zeros(n))
        for sub in range(10):
            g = {k: v.cpu().numpy() for k, v in b1.get_state().items()}
            ncon, geom, dist = b1.contacts(160)
            dbg = b1.debug_forward()
            b1.physics_step(1)
            g1 = {k: v.cpu().numpy() for k, v in b1.get_state().items()}
            for k in range(n):
                if done[k] or not np.isfinite(g["qvel"][k]).all():
                    continue                                        # auto-reset inside b.step: b1 has no task logic to follow it
                d = ref.RefData(om)
                d.qpos[:] = g["qpos"][k]; d.qvel[:] = g["qvel"][k]; d.ctrl[:] = g["ctrl"][k]; d.qacc_warmstart[:] = g["qacc_warmstart"][k]
                ref.mj_forward(om, d)
                nc = int(ncon[k]); maxcon = max(maxcon, nc)
                assert nc == d.ncon and int(dbg["nefc"][k]) == d.nefc, (s, sub, k, nc, d.ncon, int(dbg["nefc"][k]), d.nefc)
                assert np.array_equal(geom[k, :nc].cpu().numpy(), np.array([(c.geom1, c.geom2) for c in d.contact], np.int32).reshape(nc, 2))
                vmax = float(np.max(np.abs(g["qvel"][k])))
                ea = rel(dbg["qacc"][k].cpu().numpy(), d.qacc)
                d.qacc_warmstart[:] = g["qacc_warmstart"][k]
                ref.mj_step(om, d)
                if d.nwarn:
                    wild += 1; continue                             # mj_checkAcc reset on the oracle's side: compared through the counters below
                ev = rel(g1["qvel"][k], d.qvel)
                worst_a = max(worst_a, ea); worst_v = max(worst_v, ev); compared += 1
                # sane states: the single-step bound; states already flying apart (|qvel| > 1e4 rad/s): one decade more
                assert ea < (1e-3 if vmax < 1e4 else 1e-2) and ev < (1e-4 if vmax < 1e4 else 1e-3), (s, sub, k, vmax, ea, ev)
    sdict = b.stats().cpu().numpy(); s1 = b1.stats().cpu().numpy()
    print(f"full-range rollout: {compared} mj_steps compared, worst qacc {worst_a:.2e}, worst stepped qvel {worst_v:.2e}, most contacts {maxcon}, "
          f"wide passes {sdict[10] + s1[10]:.0f}, checkAcc resets on the oracle side {wild}")
    assert sdict[4] == 0 and sdict[5] == 0 and sdict[6] == 0 and s1[4] == 0 and s1[5] == 0 and s1[6] == 0
    assert compared >= 1000 and maxcon > 32
    b.close(); b1.close()


def test_state_roundtrip_autoreset_and_determinism(gpu):
    torch = gpu["torch"]
    n = 64
    def run(seed, offset, n_envs, steps=3):
        b = gpu["capi"].Batch(gpu["model"], gpu["spec"].describe(gpu["tables"]), n_envs, seed, offset)
        obs = torch.zeros((n_envs, 95), device="cuda"); rew = torch.zeros(n_envs, device="cuda")
        term = torch.zeros(n_envs, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(n_envs, dtype=torch.uint8, device="cuda")
        b.reset(obs)
        g = torch.Generator(device="cuda"); g.manual_seed(3)
        hi = torch.tensor(gpu["spec"].action_space(gpu["tables"]).high, device="cuda")
        acts = (torch.rand((steps, n, 16), device="cuda", generator=g) * 2 - 1) * hi * 0.05
        for s in range(steps):
            b.step(acts[s, offset:offset + n_envs].contiguous(), obs, rew, term, trunc)
        out = obs.clone(); st = b.get_state(); b.close()
        return out, st
    o1, s1 = run(7, 0, n)
    o2, s2 = run(7, 0, n)
    assert torch.equal(o1, o2) and torch.equal(s1["qpos"], s2["qpos"])            # deterministic
    o3, _ = run(7, 32, 32)
    assert torch.equal(o1[32:], o3)                                                 # independent of the sharding
    o4, _ = run(8, 0, n)
    assert not torch.equal(o1, o4)                                                  # seed matters (reset draws)
    # get/set state round trip
    b = gpu["capi"].Batch(gpu["model"], None, 4, 0, 0)
    q = torch.rand((4, 38)); q[:, 3:7] = torch.tensor([1.0, 0, 0, 0]); v = torch.rand((4, 37)); c = torch.rand((4, 31))
    b.set_state(q, v, c, torch.zeros((4, 37)), torch.arange(4.0))
    st = b.get_state()
    assert torch.equal(st["qpos"].cpu(), q) and torch.equal(st["qvel"].cpu(), v) and torch.equal(st["ctrl"].cpu(), c)
    assert st["time"].cpu().tolist() == [0, 1, 2, 3]
    b.close()
    # auto-reset: a fallen robot (z < 0.15) terminates, the episode is counted, the next observation is a fresh episode
    tb = gpu["capi"].Batch(gpu["model"], gpu["spec"].describe(gpu["tables"]), 2, 1, 0)
    obs = torch.zeros((2, 95), device="cuda"); fin = torch.zeros((2, 95), device="cuda"); rew = torch.zeros(2, device="cuda")
    term = torch.zeros(2, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(2, dtype=torch.uint8, device="cuda")
    tb.reset(obs)
    st = tb.get_state(); qq = st["qpos"].clone(); qq[0, 2] = 0.05
    tb.set_state(qq, st["qvel"], st["ctrl"], st["qacc_warmstart"], st["time"])
    tb.step(torch.zeros((2, 16), device="cuda"), obs, rew, term, trunc, fin)
    assert term.cpu().tolist() == [1, 0]
    assert float(fin[0, 44]) < 0.2 and abs(float(obs[0, 44]) - 0.6) < 0.05 and abs(float(obs[0, 42]) - 2.0) < 0.05
    ti, tf = tb.get_task_state()
    assert int(ti[0, 0]) == 0 and int(ti[1, 0]) == 1
    assert tb.stats().cpu().numpy()[0] == 1.0
    tb.close()


def test_host_buffer_entry_point_matches_device_path(gpu):
    torch = gpu["torch"]
    n = 32
    desc = gpu["spec"].describe(gpu["tables"])
    b1 = gpu["capi"].Batch(gpu["model"], desc, n, 3, 0); b2 = gpu["capi"].Batch(gpu["model"], desc, n, 3, 0)
    obs = torch.zeros((n, 95), device="cuda"); rew = torch.zeros(n, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(n, dtype=torch.uint8, device="cuda")
    b1.reset(obs); b2.reset(obs.clone())
    a = (np.random.default_rng(0).uniform(-1, 1, (n, 16)) * 4).astype(np.float32)
    b1.step(torch.tensor(a, device="cuda"), obs, rew, term, trunc)
    ho = np.zeros((n, 95), np.float32); hr = np.zeros(n, np.float32); ht = np.zeros(n, np.uint8); hu = np.zeros(n, np.uint8)
    b2.step_host(a, ho, hr, ht, hu)
    assert np.array_equal(ho, obs.cpu().numpy()) and np.array_equal(hr, rew.cpu().numpy()) and np.array_equal(ht, term.cpu().numpy())
    b1.close(); b2.close()


def test_vector_env_and_class_api(gpu):
    torch = gpu["torch"]
    from mujoco_gymnasium_environments_b200.vector_env import B200VectorEnv
    from mujoco_gymnasium_environments_b200.envs import QuadrupedParkourEnv
    env = B200VectorEnv("quadruped_parkour", 128, seed=1)
    obs, info = env.reset(seed=1)
    assert obs.shape == (128, 95) and obs.is_cuda and env.single_action_space.shape == (16,)
    o, r, te, tr, infos = env.step(env.action_space.sample() * 0.01)
    assert r.shape == (128,) and te.dtype == torch.bool and "final_obs" in infos
    caps = env.step_dlpack(torch.utils.dlpack.to_dlpack(torch.zeros((128, 16), device="cuda")))
    assert len(caps) == 4
    stats = env.episode_stats()
    assert stats["contacts_dropped"] == 0 and stats["substeps"] >= 128 * 30
    env.close()
    e = QuadrupedParkourEnv(render_mode=None)
    o, info = e.reset(seed=0)
    assert o.shape == (95,) and o.dtype == np.float32 and info["step_count"] == 0
    o, r, te, tr, info = e.step(np.zeros(16, np.float32))
    assert isinstance(r, float) and isinstance(te, bool) and info["step_count"] == 1
    e.close()
