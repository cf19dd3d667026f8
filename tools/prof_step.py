"""Tiny driver for ncu: N envs, reset, a few steps with random actions (scale from argv)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from mujoco_gymnasium_environments_b200.vector_env import B200VectorEnv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
task = sys.argv[4] if len(sys.argv) > 4 else "quadruped_parkour"
env = B200VectorEnv(task, N, seed=1)
env.reset()
hi = torch.tensor(env.single_action_space.high, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(0)
for i in range(steps):
    a = (torch.rand((N, hi.numel()), device="cuda", generator=g) * 2 - 1) * hi * scale
    env.step(a)
torch.cuda.synchronize()
print("ok", env.episode_stats())
