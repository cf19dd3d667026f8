"""Env sharding across GPUs: one independent shard per rank, no data-path collective.

Env i of the global batch lives on rank i // envs_per_gpu; RNG streams are keyed by the *global* env index
(``env_offset``), so results do not depend on the number of GPUs.  The only collective is the all-reduce of the
episode-statistics vector (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Sequence

STAT_KEYS = ["episodes", "return_sum", "length_sum", "nan_resets", "contacts_dropped", "rows_dropped",
             "arena_overflows", "solver_iters", "substeps", "newton_iteration_caps", "wide_passes", "wide_passes_rows"]


def shard_range(rank: int, world_size: int, envs_per_gpu: int):
    """(env_offset, n_envs) of ``rank``: weak scaling, fixed work per GPU."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return rank * envs_per_gpu, envs_per_gpu


def owner_of(env_index: int, envs_per_gpu: int) -> int:
    return env_index // envs_per_gpu


def all_reduce_stats(stats_tensor, group=None):
    """In-place SUM all-reduce of the 16-entry statistics vector when a process group exists."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(stats_tensor, op=dist.ReduceOp.SUM, group=group)
    return stats_tensor


def stats_dict(values: Sequence[float]) -> Dict[str, float]:
    out = {k: (float(values[i]) if i < len(values) else 0.0) for i, k in enumerate(STAT_KEYS)}
    n = out["episodes"]
    out["mean_return"] = out["return_sum"] / n if n else float("nan")
    out["mean_length"] = out["length_sum"] / n if n else float("nan")
    return out


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed durations are reported as the max over ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
