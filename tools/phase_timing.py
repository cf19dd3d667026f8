"""Per-phase cycle breakdown of the step kernel (needs libb2env built with -DB2_PHASE_TIMING)."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from mujoco_gymnasium_environments_b200 import capi
from mujoco_gymnasium_environments_b200.vector_env import B200VectorEnv
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
task = sys.argv[4] if len(sys.argv) > 4 else "quadruped_parkour"
preroll = int(sys.argv[5]) if len(sys.argv) > 5 else 0
env = B200VectorEnv(task, N, seed=1)
env.reset()
hi = torch.tensor(env.single_action_space.high, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(0)
for i in range(preroll):
    env.step((torch.rand((N, hi.numel()), device="cuda", generator=g) * 2 - 1) * hi * scale)
torch.cuda.synchronize()
L = capi.lib(); L.b2_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
base = (ctypes.c_ulonglong * 32)(); L.b2_phase_cycles(env.batch.handle, base)
st0 = env.episode_stats()
for i in range(steps):
    env.step((torch.rand((N, hi.numel()), device="cuda", generator=g) * 2 - 1) * hi * scale)
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 32)()
L.b2_phase_cycles(env.batch.handle, out)
out = [a - b for a, b in zip(out, base)]
names = ["kinematics+com_pos", "pass0: dynamics||contacts + factor + solve", "pass1: (J,A,PGS below) + qfc + solve", "pass2: factor(M+hD) + solve", "-", "-", "-", "-", "-",
         "fill_rows", "build_A", "pgs / newton", "wide pass (fill + solve)", "euler integrate"]
tot = sum(out[:14]) or 1
st = {k: v - st0[k] for k, v in env.episode_stats().items()}
print(task, "N", N, "scale", scale, "envs/cta", env.batch.envs_per_block, "smem", env.batch.smem_bytes); print("substeps", st["substeps"], "iters/substep", st["solver_iters"] / max(st["substeps"], 1), "wide passes", st["wide_passes"], "nan", st["nan_resets"])
nw = max(out[15], 1)
print(f"wide PGS solves {out[15]}: rows/solve {out[8]/nw:.0f} iters/solve {out[7]/nw:.1f} ring depth {out[14]/nw:.1f}; cycles/solve: B build+records {out[4]/nw:.0f}, init (v, cost) {out[5]/nw:.0f}, sweeps {out[6]/nw:.0f}")
print(f"  ring waits {out[17]} ({out[17]/nw:.0f}/solve): {out[16]/max(out[17],1):.0f} cycles each; warp-0 time at the per-iteration team barrier {out[18]/nw:.0f} cycles/solve")
if out[16]:
    print(f"  Newton islands <= 8 dofs: {out[16]} solves ({out[16]/max(st['substeps'],1):.1f}/substep), {out[17]/out[16]:.2f} iterations, {out[18]/out[16]:.0f} rows, {out[19]/out[16]:.0f} cycles each")
if out[20]:
    nb = out[20]; tb = sum(out[22:32])
    print(f"  Newton islands > 8 dofs: {nb} solves ({nb/max(st['substeps'],1):.2f}/substep), {out[21]/nb:.2f} iterations, {tb/nb:.0f} cycles each:")
    for nm, c in (("setup", 22), ("J a + gradient", 23), ("H build", 24), ("+ M, scaling", 25), ("Cholesky", 26), ("triangular solves", 27), ("M s, J s", 28), ("line search", 29), ("update, exit tests", 30), ("finish", 31)):
        print(f"     {nm:20s} {100*out[c]/tb:5.1f}%  {out[c]/nb:9.0f} cycles/solve")
if not out[15]:
    ss = max(st["substeps"], 1)
    print(f"  contact chain (the team's second warp, beside the dynamics chain): collision {out[8]/ss:.0f} (cull + slots {out[4]/ss:.0f}, narrow phase {out[5]/ss:.0f}), make_rows {out[6]/ss:.0f} cycles/substep; warp 0 waiting for it at the end of pass 0: {out[7]/ss:.0f}")
out = list(out); out[4] = out[5] = out[6] = out[7] = out[8] = 0
for n, v in zip(names, out[:14]):
    print(f"{n:32s} {100*v/tot:5.1f}%  {v/max(st['substeps'],1):10.0f} cycles/substep")
print("total cycles/substep", tot / max(st["substeps"], 1))
