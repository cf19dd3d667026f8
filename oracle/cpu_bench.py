"""CPU baseline worker: the oracle port of one task (Python task layer + C mj_step) stepped in a plain env loop.

TEST/BENCH INFRASTRUCTURE.  Mirrors how the reference would be run on host cores (one process per env, as
gymnasium.vector.AsyncVectorEnv does): each worker owns one env, draws uniform actions over the action space,
auto-resets on done, and reports (env_steps, seconds).
"""
import os
import sys
import time

import numpy as np


def worker(args):
    task, seed, n_steps, action_scale, root = args
    sys.path.insert(0, root)
    from oracle.tasks_ref import TASKS
    env = TASKS[task](seed=seed)
    rng = np.random.default_rng(seed)
    env.reset()
    lo, hi = env.action_low * action_scale, env.action_high * action_scale
    for _ in range(3):
        env.step(rng.uniform(lo, hi))
    t0 = time.perf_counter()
    for _ in range(n_steps):
        _, _, term, trunc, _ = env.step(rng.uniform(lo, hi))
        if term or trunc:
            env.reset()
    return n_steps, time.perf_counter() - t0


def run(task: str, n_procs: int, n_steps: int, action_scale: float = 1.0):
    """n_procs processes x n_steps control steps; returns (total env-steps, wall seconds, per-process seconds)."""
    import multiprocessing as mp
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(n_procs) as pool:
        res = pool.map(worker, [(task, 1000 + i, n_steps, action_scale, root) for i in range(n_procs)])
    wall = time.perf_counter() - t0
    inner = max(r[1] for r in res)
    return sum(r[0] for r in res), wall, inner
