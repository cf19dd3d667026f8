"""GPU bring-up for one task: CUDA engine vs fp64 oracle, stage by stage (prints, no asserts).

usage: python tools/bringup_task.py humanoid_dancing [action_scale]
"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch

from mujoco_gymnasium_environments_b200 import capi
from mujoco_gymnasium_environments_b200.tasks import TASKS, load_tables
from mujoco_gymnasium_environments_b200.vector_env import B200VectorEnv
from oracle import ref
from oracle.tasks_ref import TASKS as REF_TASKS


def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-12))


def make_inject(task, rng, n):
    if task == "quadruped_parkour":
        inj = np.zeros((n, 4), np.float32); inj[:, 0] = rng.uniform(-1.5, 1.5, n); inj[:, 1] = rng.uniform(-1, 1, n)
        return inj
    if task == "humanoid_dancing":
        inj = np.zeros((n, 40), np.float32)
        inj[:, 0::2] = rng.integers(0, 10, (n, 20)); inj[:, 1::2] = rng.uniform(1, 3, (n, 20))
        return inj
    if task == "humanoid_soccer":
        inj = np.zeros((n, 36), np.float32)
        inj[:, 0] = rng.uniform(-15, -5, n); inj[:, 1] = rng.uniform(-10, 10, n); inj[:, 2] = rng.uniform(-.5, .5, n)
        inj[:, 3:32] = rng.uniform(-.1, .1, (n, 29)); inj[:, 32] = rng.uniform(-2, 2, n); inj[:, 33] = rng.uniform(0, 2, n)
        inj[:, 34] = rng.uniform(0, 2 * np.pi, n); inj[:, 35] = rng.uniform(.05, .15, n)
        return inj
    if task == "bipedal_rescue":
        inj = np.zeros((n, 12), np.float32)
        inj[:, 0:2] = rng.uniform(-5, 5, (n, 2)); inj[:, 2:] = rng.uniform(-1, 1, (n, 10))
        return inj
    if task == "humanoid_construction":
        inj = np.zeros((n, 4), np.float32)
        inj[:, 0] = rng.integers(0, 4, n); inj[:, 1] = rng.uniform(0, 5, n); inj[:, 2] = rng.uniform(0, .5, n); inj[:, 3] = rng.uniform(15, 35, n)
        return inj
    if task == "humanoid_martial_arts":
        return rng.uniform(-0.5, 0.5, (n, 2)).astype(np.float32)
    if task == "robotic_arm_assembly":
        return np.zeros((n, 1), np.float32)
    raise KeyError(task)


def ref_reset(task, env, inj):
    if task == "quadruped_parkour":
        return env.reset(randomize=(float(inj[0]), float(inj[1])))
    if task == "humanoid_dancing":
        return env.reset(sequence=[(int(inj[2 * k]), float(inj[2 * k + 1])) for k in range(20)])
    if task in ("humanoid_soccer", "bipedal_rescue"):
        return env.reset(draws=[float(x) for x in inj])
    if task in ("humanoid_construction", "humanoid_martial_arts"):
        return env.reset(draws=tuple(float(x) for x in inj))
    if task == "robotic_arm_assembly":
        return env.reset()


def main():
    task = sys.argv[1] if len(sys.argv) > 1 else "humanoid_dancing"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
    spec = TASKS[task]
    t = load_tables(task)
    N = 8
    rng = np.random.default_rng(0)
    # ---------------- physics level: states sampled from an oracle rollout under random controls
    dm = capi.DeviceModel(t, 0)
    b = capi.Batch(dm, None, N, 0, 0)
    print("dims nq nv nu", b.nq, b.nv, b.nu, "smem", b.smem_bytes, "epb", b.envs_per_block, "arena", b.arena_floats,
          "ws_bytes", b.ws_bytes, "row_cap", b.row_cap, "con_cap", b.con_cap)
    renv = REF_TASKS[task](t)
    ref_reset(task, renv, make_inject(task, rng, 1)[0])
    om, od = renv.model, renv.data
    hi = spec.action_space(t).high
    states = []
    for k in range(N):
        for _ in range(5 + 7 * k):
            od.ctrl[:spec.act_dim] = rng.uniform(-1, 1, spec.act_dim) * hi * scale * (1 + k)
            ref.mj_step(om, od)
        states.append((od.qpos.copy(), od.qvel.copy(), od.ctrl.copy(), od.qacc_warmstart.copy()))
    q = torch.tensor(np.stack([s[0] for s in states]), dtype=torch.float32)
    v = torch.tensor(np.stack([s[1] for s in states]), dtype=torch.float32)
    c = torch.tensor(np.stack([s[2] for s in states]), dtype=torch.float32)
    w = torch.tensor(np.stack([s[3] for s in states]), dtype=torch.float32)
    b.set_state(q, v, c, w, torch.zeros(N))
    dbg = b.debug_forward(); torch.cuda.synchronize()
    xpos = b.xpos().cpu().numpy()
    ncon, geom, dist = b.contacts(); torch.cuda.synchronize()

    def ref_data(k):
        d = ref.RefData(om)
        d.qpos[:] = q[k].numpy().astype(np.float64); d.qvel[:] = v[k].numpy().astype(np.float64)
        d.ctrl[:] = c[k].numpy().astype(np.float64); d.qacc_warmstart[:] = w[k].numpy().astype(np.float64)
        return d

    for k in range(N):
        d = ref_data(k)
        ref.mj_forward(om, d)
        print(f"env {k}: xpos {relerr(xpos[k], d.xpos):.2e} qfrc_smooth {relerr(dbg['qfrc_smooth'][k].cpu(), d.qfrc_smooth):.2e} "
              f"qacc_smooth {relerr(dbg['qacc_smooth'][k].cpu(), d.qacc_smooth):.2e} ncon {int(dbg['ncon'][k])}/{d.ncon} "
              f"nefc {int(dbg['nefc'][k])}/{d.nefc} iters {int(dbg['solver_iter'][k])}/{d.solver_iter} "
              f"qfrc_c {relerr(dbg['qfrc_constraint'][k].cpu(), d.qfrc_constraint):.2e} qacc {relerr(dbg['qacc'][k].cpu(), d.qacc):.2e}")
        oc = [(cc.geom1, cc.geom2) for cc in d.contact]
        gc = [tuple(x) for x in geom[k, :int(ncon[k])].cpu().numpy().tolist()]
        if oc != gc:
            print("   contact pairs differ", oc, gc)
        elif oc:
            od_ = np.array([cc.dist for cc in d.contact])
            print("   contact pairs", oc, "dist err", float(np.max(np.abs(od_ - dist[k, :len(od_)].cpu().numpy()))))
    b.set_state(q, v, c, w, torch.zeros(N))
    b.physics_step(1); torch.cuda.synchronize()
    st = b.get_state()
    for k in range(N):
        d = ref_data(k)
        ref.mj_step(om, d)
        print(f"step1 env {k}: qpos {relerr(st['qpos'][k].cpu(), d.qpos):.2e} qvel {relerr(st['qvel'][k].cpu(), d.qvel):.2e} "
              f"(abs {float(np.max(np.abs(st['qvel'][k].cpu().numpy() - d.qvel))):.2e}) warm {relerr(st['qacc_warmstart'][k].cpu(), d.qacc_warmstart):.2e}")
    print("stats", b.stats().cpu().numpy()[:9])
    b.close()
    # ---------------- task level: reset with injected draws, then steps
    n = 4
    env = B200VectorEnv(task, n, device=0, seed=0)
    inj = make_inject(task, rng, n)
    obs, _ = env.reset(options={"inject": inj}); torch.cuda.synchronize()
    refs = []
    for k in range(n):
        e = REF_TASKS[task](t)
        ro, _ = ref_reset(task, e, inj[k])
        refs.append(e)
        print(f"reset env {k}: obs err {np.max(np.abs(obs[k].cpu().numpy() - ro)):.2e}")
    for s in range(int(os.environ.get("STEPS", 30))):
        act = (rng.uniform(-1, 1, (n, spec.act_dim)) * hi * scale).astype(np.float32)
        obs, rew, term, trunc, _ = env.step(act); torch.cuda.synchronize()
        for k in range(n):
            ro, rr, rt, rtr, _ = refs[k].step(act[k])
            if s % 5 == 0 or rt:
                e = np.abs(obs[k].cpu().numpy() - ro); j = int(np.argmax(e))
                print(f"step {s} env {k}: obs err {e.max():.2e} @ {j} ({float(obs[k][j]):.4f}/{ro[j]:.4f}) rew {float(rew[k]):.4f}/{rr:.4f} "
                      f"term {bool(term[k])}/{rt} ncon {refs[k].data.ncon}")
    print("stats", env.episode_stats())
    env.close()
    # ---------------- throughput
    for N in ((2048,) if task in ("bipedal_rescue", "humanoid_construction", "robotic_arm_assembly") else (4096, 8192)):
        for sc in (0.02, 1.0):
            env = B200VectorEnv(task, N, device=0, seed=1)
            env.reset()
            g = torch.Generator(device="cuda"); g.manual_seed(1)
            hi_t = torch.tensor(hi, device="cuda")
            acts = [(torch.rand((N, spec.act_dim), device="cuda", generator=g) * 2 - 1) * hi_t * sc for _ in range(8)]
            for i in range(5):
                env.step(acts[i % 8])
            torch.cuda.synchronize(); t0 = time.time()
            K = 20
            for i in range(K):
                env.step(acts[i % 8])
            torch.cuda.synchronize(); dt = (time.time() - t0) / K
            print(f"N={N} scale={sc}: {dt * 1e3:.3f} ms/step -> {N / dt:.0f} env-steps/s; stats {env.episode_stats()}")
            env.close()


if __name__ == "__main__":
    main()
