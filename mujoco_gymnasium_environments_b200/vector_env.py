"""Gymnasium-style VectorEnv over the CUDA engine: N lock-stepped envs of one task on one GPU.

Drop-in for the reference's ``env.step`` loop: ``reset(seed=...) -> (obs[N,D], infos)``,
``step(actions[N,A]) -> (obs, rew, term, trunc, infos)`` with torch CUDA tensors (exported/imported with DLPack when
the caller uses another framework) and same-step auto-reset (``infos["final_obs"]`` holds terminal observations).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import capi
from .spaces import batch_box
from .tasks import TASKS, load_tables


class B200VectorEnv:
    metadata = {"autoreset_mode": "same_step", "render_modes": []}

    def __init__(self, task: str, num_envs: int, device: int = 0, seed: int = 0, env_offset: int = 0,
                 assets_root: Optional[str] = None, **batch_opts):
        import torch
        if not torch.cuda.is_available():
            raise capi.B2Error("B200VectorEnv needs a CUDA device; there is no CPU fallback")
        self.torch = torch
        self.spec = TASKS[task]
        self.tables = load_tables(task, assets_root)
        self.model = capi.DeviceModel(self.tables, device)
        self.batch = capi.Batch(self.model, self.spec.describe(self.tables), num_envs, seed, env_offset, **batch_opts)
        self.num_envs = num_envs
        self.device = self.batch.device
        self.single_action_space = self.spec.action_space(self.tables)
        self.single_observation_space = self.spec.observation_space(self.tables)
        self.action_space = batch_box(self.single_action_space, num_envs)
        self.observation_space = batch_box(self.single_observation_space, num_envs)
        f32 = torch.float32
        self._obs = torch.empty((num_envs, self.spec.obs_dim), dtype=f32, device=self.device)
        self._final_obs = torch.zeros((num_envs, self.spec.obs_dim), dtype=f32, device=self.device)
        self._rew = torch.empty((num_envs,), dtype=f32, device=self.device)
        self._term = torch.empty((num_envs,), dtype=torch.uint8, device=self.device)
        self._trunc = torch.empty((num_envs,), dtype=torch.uint8, device=self.device)

    # ---- Gymnasium VectorEnv surface
    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        mask = inject = None
        if options:
            if options.get("reset_mask") is not None:
                mask = self.torch.as_tensor(options["reset_mask"], device=self.device).to(self.torch.uint8).contiguous()
            if options.get("inject") is not None:
                inject = self.torch.as_tensor(options["inject"], dtype=self.torch.float32, device=self.device).contiguous()
        self.batch.reset(self._obs, mask, inject)
        return self._obs, {}

    def step(self, actions):
        t = self.torch
        if not isinstance(actions, t.Tensor):
            actions = t.as_tensor(np.asarray(actions, np.float32))
        actions = actions.to(self.device, t.float32).contiguous()
        self.batch.step(actions, self._obs, self._rew, self._term, self._trunc, self._final_obs)
        term = self._term.bool(); trunc = self._trunc.bool()
        infos = {"final_obs": self._final_obs, "_final_obs": term | trunc}
        return self._obs, self._rew, term, trunc, infos

    def step_dlpack(self, actions_capsule):
        """Same as :meth:`step` for callers holding a DLPack capsule (JAX/CuPy); returns DLPack capsules."""
        from torch.utils import dlpack
        obs, rew, term, trunc, _ = self.step(dlpack.from_dlpack(actions_capsule))
        return tuple(dlpack.to_dlpack(x) for x in (obs, rew, term.to(self.torch.uint8), trunc.to(self.torch.uint8)))

    def episode_stats(self, all_reduce: bool = False):
        """Episode statistics summed over envs; with ``all_reduce`` also over ranks (NCCL, the only collective)."""
        s = self.batch.stats()
        if all_reduce:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(s, op=dist.ReduceOp.SUM)
        from .sharding import STAT_KEYS
        v = s.cpu().numpy()
        return {k: float(v[i]) for i, k in enumerate(STAT_KEYS)}

    def close(self):
        self.batch.close(); self.model.close()
