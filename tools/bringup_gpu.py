"""GPU bring-up: compare the CUDA engine with the fp64 oracle stage by stage (prints, no asserts)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch

from mujoco_gymnasium_environments_b200 import capi
from mujoco_gymnasium_environments_b200.tasks import TASKS, load_tables
from oracle import ref
from oracle.tasks_ref import QuadrupedParkourRef


def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-12))


def main():
    t = load_tables("quadruped_parkour")
    dm = capi.DeviceModel(t, 0)
    N = 8
    b = capi.Batch(dm, None, N, 0, 0)
    print("dims nq nv nu", b.nq, b.nv, b.nu, "smem", b.smem_bytes, "epb", b.envs_per_block, "arena", b.arena_floats, "ws_bytes", b.ws_bytes, "row_cap", b.row_cap, "con_cap", b.con_cap)
    om = ref.load_model(t)
    rng = np.random.default_rng(0)
    # states: oracle rollouts with small random torques
    states = []
    od = ref.RefData(om)
    od.qpos[0:3] = [2, 0, 0.6]
    for k in range(N):
        od.ctrl[:16] = rng.uniform(-1, 1, 16) * (0.5 if k < 4 else 20.0)
        ref.mj_step(om, od, 7 + 13 * k)
        states.append((od.qpos.copy(), od.qvel.copy(), od.ctrl.copy(), od.qacc_warmstart.copy()))
    q = torch.tensor(np.stack([s[0] for s in states]), dtype=torch.float32)
    v = torch.tensor(np.stack([s[1] for s in states]), dtype=torch.float32)
    c = torch.tensor(np.stack([s[2] for s in states]), dtype=torch.float32)
    w = torch.tensor(np.stack([s[3] for s in states]), dtype=torch.float32)
    b.set_state(q, v, c, w, torch.zeros(N))
    dbg = b.debug_forward(); torch.cuda.synchronize()
    xpos = b.xpos().cpu().numpy()
    ncon, geom, dist = b.contacts(); torch.cuda.synchronize()
    madr = t.dof_Madr
    for k in range(N):
        d = ref.RefData(om)
        d.qpos[:] = states[k][0]; d.qvel[:] = states[k][1]; d.ctrl[:] = states[k][2]; d.qacc_warmstart[:] = states[k][3]
        # use the fp32-rounded state the GPU saw
        d.qpos[:] = q[k].numpy().astype(np.float64); d.qvel[:] = v[k].numpy().astype(np.float64)
        d.ctrl[:] = c[k].numpy().astype(np.float64); d.qacc_warmstart[:] = w[k].numpy().astype(np.float64)
        ref.mj_forward(om, d)
        M = d.M
        Ms = np.zeros(int(t.nM))
        from mujoco_gymnasium_environments_b200.device_pack import build_device_tables
        if k == 0:
            DT = build_device_tables(t)
        for i in range(t.nv):
            a = madr[i]
            for kk in range(DT["dof_depth"][i] + 1):
                Ms[a + kk] = M[i, DT["Mcol"][a + kk]]
        print(f"env {k}: xpos {relerr(xpos[k], d.xpos):.2e} M {relerr(dbg['M'][k].cpu(), Ms):.2e} "
              f"qfrc_smooth {relerr(dbg['qfrc_smooth'][k].cpu(), d.qfrc_smooth):.2e} qacc_smooth {relerr(dbg['qacc_smooth'][k].cpu(), d.qacc_smooth):.2e} "
              f"ncon {int(dbg['ncon'][k])}/{d.ncon} nefc {int(dbg['nefc'][k])}/{d.nefc} iters {int(dbg['solver_iter'][k])}/{d.solver_iter} "
              f"qfrc_c {relerr(dbg['qfrc_constraint'][k].cpu(), d.qfrc_constraint):.2e} qacc {relerr(dbg['qacc'][k].cpu(), d.qacc):.2e}")
        oc = [(cc.geom1, cc.geom2) for cc in d.contact]
        gc = [tuple(x) for x in geom[k, :int(ncon[k])].cpu().numpy().tolist()]
        if oc != gc:
            print("   contact pairs differ", oc, gc)
        else:
            od_ = np.array([cc.dist for cc in d.contact]);
            if len(od_): print("   contact dist err", float(np.max(np.abs(od_ - dist[k, :len(od_)].cpu().numpy()))))
    # one physics step
    b.set_state(q, v, c, w, torch.zeros(N))
    b.physics_step(1); torch.cuda.synchronize()
    st = b.get_state()
    for k in range(N):
        d = ref.RefData(om)
        d.qpos[:] = q[k].numpy().astype(np.float64); d.qvel[:] = v[k].numpy().astype(np.float64)
        d.ctrl[:] = c[k].numpy().astype(np.float64); d.qacc_warmstart[:] = w[k].numpy().astype(np.float64)
        ref.mj_step(om, d)
        print(f"step1 env {k}: qpos {relerr(st['qpos'][k].cpu(), d.qpos):.2e} qvel {relerr(st['qvel'][k].cpu(), d.qvel):.2e} "
              f"(abs {float(np.max(np.abs(st['qvel'][k].cpu().numpy() - d.qvel))):.2e}) warm {relerr(st['qacc_warmstart'][k].cpu(), d.qacc_warmstart):.2e}")
    # 100 physics steps drift
    b.set_state(q, v, c, w, torch.zeros(N))
    b.physics_step(100); torch.cuda.synchronize()
    st = b.get_state()
    for k in range(N):
        d = ref.RefData(om)
        d.qpos[:] = q[k].numpy().astype(np.float64); d.qvel[:] = v[k].numpy().astype(np.float64)
        d.ctrl[:] = c[k].numpy().astype(np.float64); d.qacc_warmstart[:] = w[k].numpy().astype(np.float64)
        ref.mj_step(om, d, 100)
        print(f"step100 env {k}: qpos abs {float(np.max(np.abs(st['qpos'][k].cpu().numpy() - d.qpos))):.2e} qvel abs {float(np.max(np.abs(st['qvel'][k].cpu().numpy() - d.qvel))):.2e}")
    print("stats", b.stats().cpu().numpy())
    b.close()

    # task-level
    spec = TASKS["quadruped_parkour"]
    N = 4
    tb = capi.Batch(dm, spec.describe(t), N, 1234, 0)
    obs = torch.zeros((N, 95), device="cuda"); rew = torch.zeros(N, device="cuda")
    term = torch.zeros(N, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(N, dtype=torch.uint8, device="cuda")
    inject = torch.tensor([[0.3, -0.2, 0, 0], [1.0, 0.5, 0, 0], [-1.2, 0.9, 0, 0], [0.0, 0.0, 0, 0]], dtype=torch.float32, device="cuda")
    tb.reset(obs, None, inject); torch.cuda.synchronize()
    envs = [QuadrupedParkourRef(t) for _ in range(N)]
    for k in range(N):
        o, _ = envs[k].reset(randomize=(float(inject[k, 0]), float(inject[k, 1])))
        print(f"reset env {k}: obs err {float(np.max(np.abs(o - obs[k].cpu().numpy()))):.2e}")
    arng = np.random.default_rng(5)
    for s in range(5):
        a = (arng.uniform(-1, 1, (N, 16)) * 0.02 * np.array(spec.action_space(t).high)).astype(np.float32)
        tb.step(torch.tensor(a, device="cuda"), obs, rew, term, trunc); torch.cuda.synchronize()
        for k in range(N):
            o, r, te, tr, info = envs[k].step(a[k])
            print(f"step {s} env {k}: obs err {float(np.max(np.abs(o - obs[k].cpu().numpy()))):.2e} rew {r:.4f}/{float(rew[k]):.4f} term {te}/{int(term[k])} ncon {envs[k].data.ncon}")
    # timing
    for N in (4096,):
        tb2 = capi.Batch(dm, spec.describe(t), N, 1, 0)
        obs = torch.zeros((N, 95), device="cuda"); rew = torch.zeros(N, device="cuda")
        term = torch.zeros(N, dtype=torch.uint8, device="cuda"); trunc = torch.zeros(N, dtype=torch.uint8, device="cuda")
        tb2.reset(obs); torch.cuda.synchronize()
        hi = torch.tensor(spec.action_space(t).high, device="cuda")
        for scale in (0.02, 1.0):
            for it in range(3):
                a = (torch.rand((N, 16), device="cuda") * 2 - 1) * hi * scale
                tb2.step(a, obs, rew, term, trunc)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            K = 10
            for it in range(K):
                a = (torch.rand((N, 16), device="cuda") * 2 - 1) * hi * scale
                tb2.step(a, obs, rew, term, trunc)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / K
            print(f"N={N} epb={tb2.envs_per_block} scale={scale}: {ms:.3f} ms/step -> {N / ms * 1e3:.0f} env-steps/s; stats {tb2.stats().cpu().numpy()[:9]}")
        tb2.close()


if __name__ == "__main__":
    main()
