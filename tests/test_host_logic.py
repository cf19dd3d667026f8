"""Host-side logic that needs no GPU: task registry, spaces, sharding, and the 2-rank gloo stats all-reduce."""
import os
import subprocess
import sys

import numpy as np
import pytest

from mujoco_gymnasium_environments_b200 import capi, sharding
from mujoco_gymnasium_environments_b200.tasks import TASKS


def test_task_registry(quad_tables):
    spec = TASKS["quadruped_parkour"]
    assert (spec.obs_dim, spec.act_dim, spec.frame_skip, spec.max_episode_steps) == (95, 16, 10, 6000)
    assert spec.bytes_per_env_step == 1426          # 4*[(38+37+37+16+10)+(38+37+37+95+1+10)]+2
    a = spec.action_space(quad_tables)
    assert a.shape == (16,) and a.high.tolist() == [80, 80, 60, 40] * 4 and np.array_equal(a.low, -a.high)
    o = spec.observation_space(quad_tables)
    assert o.shape == (95,) and o.low[0] == pytest.approx(-np.pi) and o.high[16] == 20 and np.isinf(o.high[42])
    d = spec.describe(quad_tables)
    assert d.task == capi.TASK_QUADRUPED_PARKOUR and list(d.ids[:9]) == [1, 5, 9, 13, 17, 17, 18, 16, 17]
    assert d.act_hi[2] == 60 and d.act_lo[3] == -40


def test_sharding_partition():
    assert sharding.shard_range(0, 4, 4096) == (0, 4096) and sharding.shard_range(3, 4, 4096) == (12288, 4096)
    assert sharding.owner_of(8191, 4096) == 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 4, 1)
    s = sharding.stats_dict([2, 10.0, 30] + [0] * 7)
    assert s["mean_return"] == 5.0 and s["mean_length"] == 15.0


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from mujoco_gymnasium_environments_b200 import sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
off, n = sharding.shard_range(rank, 2, 8)
stats = torch.zeros(16, dtype=torch.float64); stats[0] = rank + 1; stats[1] = 10.0 * (rank + 1); stats[2] = off
sharding.all_reduce_stats(stats)
t = sharding.max_over_ranks(1.5 + rank)
assert stats[0].item() == 3 and stats[1].item() == 30.0 and stats[2].item() == 8 and t == 2.5, (stats, t)
dist.barrier(); dist.destroy_process_group()
print("OK", rank)
'''


def test_two_rank_gloo_stats_allreduce(tmp_path):
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    script = tmp_path / "w.py"; script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), root, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("OK" in o for o in outs)
