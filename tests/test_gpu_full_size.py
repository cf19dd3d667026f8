"""Properties checked at BASELINE.json's full batch sizes (4096 / 8192 / 2048 envs per GPU), where a per-env comparison
against the oracle would take minutes:
  * lockstep uniformity: every env reset with the same injected draws and driven by the same actions must produce
    bit-identical observation / reward / flag rows, whatever CTA, team slot or wave it runs in -- and row 0 must match
    the fp64 oracle (one oracle env stands for all N);
  * shard invariance: a slice [o, o+8) of the full batch run as its own batch with env_offset = o and the kernel's own
    RNG reproduces the same rows bit for bit (what makes multi-GPU sharding transparent);
  * determinism: a second run of the same batch is bit-identical.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = {"quadruped_parkour": 4096, "humanoid_dancing": 8192, "humanoid_soccer": 4096, "bipedal_rescue": 2048,
         "humanoid_construction": 2048, "humanoid_martial_arts": 4096, "robotic_arm_assembly": 2048}


def _inject(task, rng):
    if task == "quadruped_parkour":
        return np.array([0.4, -0.3, 0, 0], np.float32)
    if task == "humanoid_dancing":
        x = np.zeros(40, np.float32); x[0::2] = rng.integers(0, 10, 20); x[1::2] = rng.uniform(1, 3, 20); return x
    if task == "humanoid_soccer":
        x = np.zeros(36, np.float32); x[0] = -8.0; x[1] = 2.0; x[2] = 0.2; x[3:32] = rng.uniform(-.1, .1, 29); x[32] = 0.5; x[33] = 1.0; x[34] = 1.0; x[35] = 0.1; return x
    if task == "bipedal_rescue":
        x = np.zeros(12, np.float32); x[:2] = [1.5, -2.5]; x[2:] = rng.uniform(-1, 1, 10); return x
    if task == "humanoid_construction":
        return np.array([0, 2.0, 0.1, 20.0], np.float32)
    if task == "humanoid_martial_arts":
        return np.array([0.3, -0.2], np.float32)
    return np.zeros(1, np.float32)


def _ref_reset(task, env, inj):
    if task == "quadruped_parkour":
        return env.reset(randomize=(float(inj[0]), float(inj[1])))
    if task == "humanoid_dancing":
        return env.reset(sequence=[(int(inj[2 * k]), float(inj[2 * k + 1])) for k in range(20)])
    if task in ("humanoid_construction", "humanoid_martial_arts"):
        return env.reset(draws=tuple(float(v) for v in inj))
    if task == "robotic_arm_assembly":
        return env.reset()
    return env.reset(draws=[float(v) for v in inj])


@pytest.mark.parametrize("task", list(SIZES))
def test_lockstep_uniformity_at_full_size(task):
    import torch
    from mujoco_gymnasium_environments_b200.vector_env import B200VectorEnv
    from oracle.tasks_ref import TASKS as REF
    n = SIZES[task]
    rng = np.random.default_rng(11)
    env = B200VectorEnv(task, n, device=0, seed=3)
    inj = _inject(task, rng)
    obs, _ = env.reset(options={"inject": np.tile(inj, (n, 1))})
    assert torch.equal(obs, obs[0:1].expand_as(obs))
    st = {k: v[0].cpu().numpy() for k, v in env.batch.get_state().items()}
    e = REF[task](env.tables)
    _ref_reset(task, e, inj)
    d = e.data
    d.qpos[:] = st["qpos"]; d.qvel[:] = st["qvel"]; d.qacc_warmstart[:] = st["qacc_warmstart"]
    if task == "humanoid_dancing":
        e.prev_joint_vel = d.qvel[6:].copy()
    hi = env.single_action_space.high
    twins = []
    if task == "robotic_arm_assembly":      # compared up to the oracle's own response to an fp32-sized perturbation (oracle/twin.py)
        from oracle.twin import SLACK, perturbed
        prng = np.random.default_rng(2)
        for _ in range(12):      # the second control step is 20 `mj_step`s into a chaotic scene: a small ensemble underestimates the bundle
            g = REF[task](env.tables); _ref_reset(task, g, inj)
            g.data.qpos[:] = perturbed(st["qpos"], prng); g.data.qvel[:] = perturbed(st["qvel"], prng); g.data.qacc_warmstart[:] = st["qacc_warmstart"]
            twins.append(g)
    for s in range(2):
        a = (rng.uniform(-1, 1, env.spec.act_dim) * hi * 0.02).astype(np.float32)
        obs, rew, term, trunc, _ = env.step(np.tile(a, (n, 1)))
        assert torch.equal(obs, obs[0:1].expand_as(obs)), (task, s)                    # all N rows bit-identical
        assert torch.equal(rew, rew[0:1].expand_as(rew)) or bool(torch.isinf(rew).all())
        assert torch.equal(term, term[0:1].expand_as(term)) and torch.equal(trunc, trunc[0:1].expand_as(trunc))
        ro, rr, rt, rtr, _ = e.step(a)
        tol = 2e-2 if task == "robotic_arm_assembly" else 2e-3
        slack = 0.0
        for g in twins:
            slack = np.maximum(slack, SLACK * np.abs(g.step(a)[0] - ro))
        assert float(np.max(np.maximum(np.abs(obs[0].cpu().numpy() - ro) - slack, 0.0) / (1.0 + np.abs(ro)))) < tol, (task, s)
        assert bool(term[0]) == rt and bool(trunc[0]) == rtr
    env.close()


@pytest.mark.parametrize("task", ["quadruped_parkour", "humanoid_construction", "humanoid_martial_arts", "robotic_arm_assembly"])
def test_shard_invariance_and_determinism_at_full_size(task):
    import torch
    from mujoco_gymnasium_environments_b200.vector_env import B200VectorEnv
    n = SIZES[task]; off = n - 13

    def run(num, env_offset):
        env = B200VectorEnv(task, num, device=0, seed=5, env_offset=env_offset)
        o0, _ = env.reset()
        hi = torch.tensor(env.single_action_space.high, device="cuda")
        outs = [o0.clone()]
        for s in range(3):
            g = torch.Generator(device="cuda"); g.manual_seed(100 + s)
            a_all = (torch.rand((n, hi.numel()), device="cuda", generator=g) * 2 - 1) * hi * 0.05
            o, r, te, tr, _ = env.step(a_all[env_offset:env_offset + num].contiguous())
            outs += [o.clone(), r.clone(), te.clone().float(), tr.clone().float()]
        env.close()
        return outs

    full = run(n, 0)
    again = run(n, 0)
    shard = run(8, off)
    for a, b in zip(full, again):
        assert torch.equal(a, b)                                                        # determinism
    for a, b in zip(full, shard):
        assert torch.equal(a[off:off + 8], b), task                                     # shard == slice of the full batch
