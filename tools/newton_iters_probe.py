"""GPU probe: per-step Newton iteration counts of humanoid_construction; dumps the states whose solve ran long."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from mujoco_gymnasium_environments_b200.tasks import TASKS, load_tables
from mujoco_gymnasium_environments_b200.vector_env import B200VectorEnv

task = "humanoid_construction"; spec = TASKS[task]; t = load_tables(task)
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.03
n = 16
env = B200VectorEnv(task, n, device=0, seed=0)
env.reset()
rng = np.random.default_rng(0); hi = spec.action_space(t).high
dump = []
prev = 0.0
for s in range(40):
    st = {k: v.cpu().numpy().copy() for k, v in env.batch.get_state().items()}
    act = (rng.uniform(-1, 1, (n, spec.act_dim)) * hi * scale).astype(np.float32)
    env.step(act); torch.cuda.synchronize()
    it = float(env.episode_stats()["solver_iters"])
    d = it - prev; prev = it
    print(s, d / n, "capped", float(env.batch.stats()[9]))
    if d / n > 6 and len(dump) < 6:
        dump.append(dict(step=s, iters=d, act=act, **st))
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/newton_long.npz", **{f"{i}_{k}": np.asarray(v) for i, dd in enumerate(dump) for k, v in dd.items()})
print("dumped", len(dump))
