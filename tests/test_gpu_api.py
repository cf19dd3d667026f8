"""The callers' side of env.step on the GPU: seeding, argument checks, CUDA-graph capture / b2_rollout, the terminal-step
snapshot behind ``info``, the pinned host buffers of b2_step_host, and a second device-agnostic batch of the same task."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def make(task, n, **kw):
    from mujoco_gymnasium_environments_b200.vector_env import B200VectorEnv
    return B200VectorEnv(task, n, device=0, **kw)


def test_reset_seed_is_honoured(torch):
    """reset(seed=s) twice -> identical initial states; different seeds -> different ones (the reference reseeds np_random,
    quadruped_parkour_env/parkour_env.py:314-322)."""
    env = make("humanoid_soccer", 16, seed=5)
    o1 = env.reset(seed=11)[0].clone(); env.step(torch.zeros((16, 33), device="cuda"))
    o2 = env.reset(seed=11)[0].clone()
    o3 = env.reset(seed=12)[0].clone()
    assert torch.equal(o1, o2)
    assert not torch.equal(o1, o3)
    o4 = env.reset()[0].clone()                     # no seed: the stream continues (next episode ids)
    assert not torch.equal(o4, o3)
    env.close()


def test_wrong_shapes_raise_instead_of_reading_out_of_bounds(torch):
    env = make("quadruped_parkour", 8)
    env.reset()
    with pytest.raises(ValueError):
        env.step(torch.zeros((8, 15), device="cuda"))
    with pytest.raises(ValueError):
        env.step(torch.zeros((16,), device="cuda"))
    with pytest.raises(ValueError):
        env.reset(options={"reset_mask": torch.ones(7, dtype=torch.uint8, device="cuda")})
    with pytest.raises(ValueError):
        env.reset(options={"inject": torch.zeros((8, 3), device="cuda")})
    assert env.batch.ninj == 4
    with pytest.raises(ValueError):
        env.batch.step(torch.zeros((8, 16)), env._obs, env._rew, env._term, env._trunc)          # CPU tensor
    ho = np.zeros((8, 95), np.float32)
    with pytest.raises(ValueError):
        env.batch.step_host(np.zeros((8, 16), np.float64), ho, np.zeros(8, np.float32), np.zeros(8, np.uint8), np.zeros(8, np.uint8))
    env.close()


@pytest.mark.parametrize("task", ["quadruped_parkour", "humanoid_dancing"])
def test_graph_replay_and_rollout_are_bit_identical_to_eager(torch, task):
    """32 steps: eager b2_step loop == torch.cuda.graph replay of the same loop == b2_rollout (one native CUDA graph)."""
    T, n = 32, 24
    envs = [make(task, n, seed=3) for _ in range(3)]
    A = envs[0].spec.act_dim; D = envs[0].spec.obs_dim
    hi = torch.tensor(envs[0].single_action_space.high, device="cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    acts = ((torch.rand((T, n, A), device="cuda", generator=g) * 2 - 1) * hi * 0.3).contiguous()
    for e in envs:
        e.reset(seed=3)
    outs = []
    for e in envs:
        outs.append(dict(obs=torch.zeros((T, n, D), device="cuda"), rew=torch.zeros((T, n), device="cuda"),
                         term=torch.zeros((T, n), dtype=torch.uint8, device="cuda"), trunc=torch.zeros((T, n), dtype=torch.uint8, device="cuda")))
    # eager
    e, o = envs[0], outs[0]
    for t in range(T):
        e.batch.step(acts[t], o["obs"][t], o["rew"][t], o["term"][t], o["trunc"][t])
    # torch CUDA graph around the same calls
    e, o = envs[1], outs[1]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(graph, stream=s):
        for t in range(T):
            e.batch.step(acts[t], o["obs"][t], o["rew"][t], o["term"][t], o["trunc"][t])
    e.reset(seed=3)                                  # capture did not execute anything; start from the same state
    torch.cuda.synchronize()
    graph.replay()
    # native rollout
    e, o = envs[2], outs[2]
    e.batch.rollout(acts, o["obs"], o["rew"], o["term"], o["trunc"])
    torch.cuda.synchronize()
    bits = lambda x: x.view(torch.int32) if x.dtype == torch.float32 else x        # bit-for-bit, NaN rewards of blown-up dancers included
    for k in ("obs", "rew", "term", "trunc"):
        assert torch.equal(bits(outs[0][k]), bits(outs[1][k])), (task, "graph", k)
        assert torch.equal(bits(outs[0][k]), bits(outs[2][k])), (task, "rollout", k)
    # a second rollout replays the cached graph and continues the episodes
    e.batch.rollout(acts, o["obs"], o["rew"], o["term"], o["trunc"])
    torch.cuda.synchronize()
    assert torch.isfinite(o["obs"]).all()
    for e in envs:
        e.close()


def test_terminal_step_info_reports_the_finished_episode(torch):
    """On a terminal step the kernel has already auto-reset; info must still carry the finished episode's totals
    (parkour_env.py:797-813), which come from the snapshot taken before the in-kernel reset."""
    from mujoco_gymnasium_environments_b200.envs import QuadrupedParkourEnv
    env = QuadrupedParkourEnv()
    env.reset(seed=1)
    rng = np.random.default_rng(0); total = 0.0; steps = 0
    for _ in range(4000):
        o, r, term, trunc, info = env.step(rng.uniform(-1, 1, 16).astype(np.float32) * env.action_space.high)
        total += r; steps += 1
        assert info["step_count"] == steps
        if term or trunc:
            break
    assert term or trunc, "full-range torques should end the episode within 4000 steps"
    assert info["step_count"] == steps and info["step_count"] > 0
    assert info["episode_reward"] == pytest.approx(total, rel=1e-4, abs=1e-2)
    o, r, term, trunc, info = env.step(np.zeros(16, np.float32))
    assert info["step_count"] == 1                                  # the next episode counts from its own start
    env.close()
    # vector API: lazily built per-env infos with the reference keys, terminal rows from the snapshot
    venv = make("quadruped_parkour", 64, seed=2)
    venv.reset()
    hi = torch.tensor(venv.single_action_space.high, device="cuda")
    seen = False
    for _ in range(200):
        obs, rew, term, trunc, infos = venv.step((torch.rand((64, 16), device="cuda") * 2 - 1) * hi)
        done = infos["_final_obs"]
        if bool(done.any()):
            sc = infos["step_count"]
            assert set(venv.spec.info_keys) <= set(infos.keys())
            assert bool((sc[done] > 0).all())                       # finished episodes report their length, not 0
            seen = True
            break
    assert seen
    venv.close()


def test_step_host_with_the_pinned_buffers_matches_device_step(torch):
    a = make("humanoid_soccer", 32, seed=4); b = make("humanoid_soccer", 32, seed=4)
    a.reset(seed=4); b.reset(seed=4)
    act, obs, rew, term, trunc = b.batch.host_buffers()
    rng = np.random.default_rng(1)
    for _ in range(5):
        x = rng.uniform(-15, 15, (32, 33)).astype(np.float32)
        o, r, te, tr, _ = a.step(torch.tensor(x, device="cuda"))
        act[:] = x
        b.batch.step_host(act, obs, rew, term, trunc)               # zero-copy: the library's own pinned staging
        assert np.array_equal(o.cpu().numpy(), obs) and np.array_equal(r.cpu().numpy(), rew)
        assert np.array_equal(te.cpu().numpy().astype(np.uint8), term)
        # pageable caller buffers take the staging copies and give the same numbers
    ho = np.zeros((32, 80), np.float32); hr = np.zeros(32, np.float32); ht = np.zeros(32, np.uint8); hu = np.zeros(32, np.uint8)
    x = rng.uniform(-15, 15, (32, 33)).astype(np.float32)
    o, r, te, tr, _ = a.step(torch.tensor(x, device="cuda"))
    b.batch.step_host(x, ho, hr, ht, hu)
    assert np.array_equal(o.cpu().numpy(), ho) and np.array_equal(r.cpu().numpy(), hr)
    a.close(); b.close()


def test_mixed_streams_are_ordered(torch):
    """b2_step_host runs on the library's own stream, reset / step on the caller's: the library orders them with an event."""
    e = make("quadruped_parkour", 256, seed=9); f = make("quadruped_parkour", 256, seed=9)
    side = torch.cuda.Stream()
    act = np.random.default_rng(2).uniform(-1, 1, (256, 16)).astype(np.float32)
    ho = np.zeros((256, 95), np.float32); hr = np.zeros(256, np.float32); ht = np.zeros(256, np.uint8); hu = np.zeros(256, np.uint8)
    with torch.cuda.stream(side):
        e.reset(seed=9)
        e.batch.step_host(act, ho, hr, ht, hu)          # must see the reset issued on `side` just before
    f.reset(seed=9); o, r, _, _, _ = f.step(torch.tensor(act, device="cuda"))
    torch.cuda.synchronize()
    assert np.array_equal(o.cpu().numpy(), ho)
    e.close(); f.close()


def test_async_pairs_and_get_attr_follow_the_device_state(torch):
    """step_async / step_wait give the same trajectory as step (two batches from one seed), and get_attr reads the per-env
    quantities the reference keeps as attributes (quadruped_parkour_env/parkour_env.py:797-813) from the device."""
    a = make("quadruped_parkour", 12, seed=3); b = make("quadruped_parkour", 12, seed=3)
    a.reset(seed=9); b.reset_async(seed=9); b.reset_wait()
    g = torch.Generator(device="cuda").manual_seed(1)
    for k in range(5):
        act = (torch.rand((12, 16), device="cuda", generator=g) - 0.5) * 4.0
        oa, ra, ta, ua, _ = a.step(act)
        b.step_async(act); ob, rb, tb, ub, _ = b.step_wait()
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(ta, tb) and torch.equal(ua, ub)
    assert b.get_attr("step_count") == (5,) * 12 and b.get_attr("max_episode_steps") == (6000,) * 12
    er = b.get_attr("episode_reward")
    assert len(er) == 12 and all(isinstance(v, float) for v in er)
    with pytest.raises(AttributeError):
        b.get_attr("np_random")
    a.close(); b.close()


def test_rescue_class_exposes_the_victim_lists_the_reference_test_reads():
    # bipedal_rescue_env/test_rescue.py:298-300 reads env.victims_rescued / victims_carried / victim_priorities
    from mujoco_gymnasium_environments_b200.envs import BipedalRescueEnv
    env = BipedalRescueEnv()
    env.reset(seed=0)
    assert env.victims_rescued == [] and env.victims_carried == [] and len(env.victim_priorities) == env.num_victims == 5
    ti, tf = env._vec.batch.get_task_state()
    ti[0, 1] = 0b00010; ti[0, 2] = 0b01001
    env._vec.batch.set_task_state(ti, tf)
    assert env.victims_rescued == [1] and env.victims_carried == [0, 3]
    env.close()


def test_set_attr_writes_task_state_columns_the_kernel_then_uses(torch):
    """set_attr("step_count", ...) lands in the device task state: the episode is truncated when the counter reaches the
    reference's max_episode_steps (the test is made before the increment, b2_tasks.cuh QuadrupedTask): on the fifth step after 5 996."""
    env = make("quadruped_parkour", 8, seed=2)
    env.reset(seed=2)
    env.set_attr("step_count", 5996)
    assert env.get_attr("step_count") == (5996,) * 8
    env.set_attr("episode_reward", [float(i) for i in range(8)])
    zero = torch.zeros((8, 16), device="cuda")
    for k in range(5):
        _, rew, term, trunc, _ = env.step(zero)
        assert not bool(term.any())
        assert bool(trunc.all()) == (k == 4), k
    assert env.get_attr("step_count") == (6001,) * 8          # the finished episode's value on the terminal step
    with pytest.raises(AttributeError):
        env.set_attr("course_completion", 0.5)                # derived from xpos
    env.close()
