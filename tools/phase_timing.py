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
env = B200VectorEnv(task, N, seed=1)
env.reset()
hi = torch.tensor(env.single_action_space.high, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(0)
for i in range(steps):
    env.step((torch.rand((N, hi.numel()), device="cuda", generator=g) * 2 - 1) * hi * scale)
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 16)()
L = capi.lib(); L.b2_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
L.b2_phase_cycles(env.batch.handle, out)
names = ["kinematics+com_pos", "pass0: dynamics||contacts + factor + solve", "pass1: (J,A,PGS below) + qfc + solve", "pass2: factor(M+hD) + solve", "-", "-", "-", "-", "-",
         "fill_rows", "build_A", "pgs", "finish", "euler(factor+solve+integrate)"]
tot = sum(out[:14]) or 1
st = env.episode_stats()
print(task, "N", N, "scale", scale, "envs/cta", env.batch.envs_per_block, "smem", env.batch.smem_bytes); print("substeps", st["substeps"], "iters/substep", st["solver_iters"] / max(st["substeps"], 1))
for n, v in zip(names, out[:14]):
    print(f"{n:32s} {100*v/tot:5.1f}%  {v/max(st['substeps'],1):10.0f} cycles/substep")
print("total cycles/substep", tot / max(st["substeps"], 1))
