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


class LazyInfos(dict):
    """``infos`` of a vector step.  ``final_obs`` / ``_final_obs`` are there from the start; the per-env counterparts of the
    reference's ``info`` dict (e.g. quadruped_parkour_env/parkour_env.py:797-813: step_count, episode_reward, fall_count ...) are
    device tensors of shape [N] built on first access from the task state -- rows of envs that finished an episode in this
    step come from the snapshot taken before the same-step auto-reset -- so a training loop that never reads them pays nothing."""

    def __init__(self, env, eager):
        super().__init__(eager)
        self._env = env; self._built = False

    def _build(self):
        if not self._built:
            self._built = True
            ti, tf, xpos = self._env.task_state(with_xpos=True)
            for k, v in self._env.spec.vector_info(self._env.torch, ti, tf, xpos, self._env.tables).items():
                dict.setdefault(self, k, v)
            dict.setdefault(self, "task_ti", ti); dict.setdefault(self, "task_tf", tf)

    def __missing__(self, key):
        self._build()
        if not dict.__contains__(self, key):
            raise KeyError(key)
        return dict.__getitem__(self, key)

    def __contains__(self, key):
        if dict.__contains__(self, key):
            return True
        self._build()
        return dict.__contains__(self, key)

    def keys(self):
        self._build()
        return dict.keys(self)


class B200VectorEnv:
    metadata = {"autoreset_mode": "same_step", "render_modes": []}

    def __init__(self, task: str, num_envs: int, device: int = 0, seed: int = 0, env_offset: int = 0,
                 assets_root: Optional[str] = None, tables=None, **batch_opts):
        import torch
        if not torch.cuda.is_available():
            raise capi.B2Error("B200VectorEnv needs a CUDA device; there is no CPU fallback")
        self.torch = torch
        self.spec = TASKS[task]
        # tables: a ModelTables made elsewhere, e.g. from_mjmodel(mujoco.MjModel.from_xml_string(...)) where MuJoCo is importable
        self.tables = tables if tables is not None else load_tables(task, assets_root)
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
        self._done = torch.zeros((num_envs,), dtype=torch.bool, device=self.device)
        self._pending = None

    # ---- Gymnasium VectorEnv surface
    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        mask = inject = None
        if seed is not None:
            # Env.reset(seed=...) reseeds np_random in the reference (parkour_env.py:314-322); here it replaces the device RNG seed and
            # restarts the episode counters, so the same seed always yields the same initial states
            self.batch.reseed(int(seed[0]) if isinstance(seed, (list, tuple)) else int(seed))
        if options:
            if options.get("reset_mask") is not None:
                mask = self.torch.as_tensor(options["reset_mask"], device=self.device).to(self.torch.uint8).contiguous()
            if options.get("inject") is not None:
                inject = self.torch.as_tensor(options["inject"], dtype=self.torch.float32, device=self.device).contiguous()
        if mask is not None and mask.numel() != self.num_envs:
            raise ValueError(f"reset_mask: expected {self.num_envs} entries, got {mask.numel()}")
        if inject is not None and tuple(inject.shape) != (self.num_envs, self.batch.ninj):
            raise ValueError(f"inject: expected shape {(self.num_envs, self.batch.ninj)}, got {tuple(inject.shape)}")
        self.batch.reset(self._obs, None if mask is None else mask.reshape(self.num_envs), inject)
        if mask is None:
            self._done.zero_()
        else:
            self._done &= ~mask.reshape(self.num_envs).bool()
        return self._obs, LazyInfos(self, {})

    def step(self, actions):
        t = self.torch
        if not isinstance(actions, t.Tensor):
            actions = t.as_tensor(np.asarray(actions, np.float32))
        if tuple(actions.shape) != (self.num_envs, self.spec.act_dim):
            raise ValueError(f"actions: expected shape {(self.num_envs, self.spec.act_dim)}, got {tuple(actions.shape)}")
        actions = actions.to(self.device, t.float32).contiguous()
        self.batch.step(actions, self._obs, self._rew, self._term, self._trunc, self._final_obs)
        term = self._term.bool(); trunc = self._trunc.bool()
        self._done = term | trunc
        infos = LazyInfos(self, {"final_obs": self._final_obs, "_final_obs": self._done})
        return self._obs, self._rew, term, trunc, infos

    # ---- the rest of gymnasium.vector.VectorEnv / AsyncVectorEnv's surface, for callers written against the reference's envs
    # under AsyncVectorEnv (BASELINE.md section 3 harness B): step_async really is asynchronous here -- it enqueues the step on
    # the batch's stream and returns -- and step_wait hands out the tensors without a host synchronisation
    def step_async(self, actions):
        if self._pending is not None:
            raise RuntimeError(f"step_async called while a {self._pending[0]} is pending")
        self._pending = ("step", self.step(actions))

    def step_wait(self, timeout=None):
        return self._take("step")

    def reset_async(self, seed=None, options=None):
        if self._pending is not None:
            raise RuntimeError(f"reset_async called while a {self._pending[0]} is pending")
        self._pending = ("reset", self.reset(seed=seed, options=options))

    def reset_wait(self, timeout=None):
        return self._take("reset")

    def _take(self, kind):
        if self._pending is None or self._pending[0] != kind:
            raise RuntimeError(f"{kind}_wait called without a pending {kind}_async")
        out = self._pending[1]; self._pending = None
        return out

    def get_attr(self, name: str):
        """Tuple of one value per env, as ``AsyncVectorEnv.get_attr``: constants of the env class (``max_episode_steps``,
        ``frame_skip``, ``render_mode``, the single-env spaces, ``metadata``) repeated, and the per-env quantities the reference keeps as
        attributes and reports in ``info`` (``step_count``, ``episode_stats`` ...) read back from the device task state."""
        consts = {"max_episode_steps": self.spec.max_episode_steps, "frame_skip": self.spec.frame_skip, "render_mode": None,
                  "action_space": self.single_action_space, "observation_space": self.single_observation_space,
                  "metadata": {"render_modes": [], "render_fps": self.spec.render_fps}}
        if name in consts:
            return (consts[name],) * self.num_envs
        if name in self.spec.info_keys:
            infos = LazyInfos(self, {})
            if name in infos:
                v = infos[name]
                if isinstance(v, dict):
                    cols = {k: c.cpu().tolist() for k, c in v.items()}
                    return tuple({k: c[i] for k, c in cols.items()} for i in range(self.num_envs))
                return tuple(v.cpu().tolist())
        raise AttributeError(f"{self.spec.name}: no per-env attribute {name!r} (constants: {sorted(consts)}; per-env: {self.spec.info_keys})")

    def set_attr(self, name: str, values):
        """``AsyncVectorEnv.set_attr`` for the per-env quantities that ARE task-state columns (``step_count``, ``episode_reward``,
        ``energy_remaining``, the leaves of ``episode_stats`` ...): one value (or dict of values) per env, or a single one for all.
        Derived quantities (``course_completion``, ``victims_remaining`` ...) and class constants cannot be set; the physics state
        is written with ``batch.set_state``."""
        t = self.torch
        ti, tf = self.batch.get_task_state()
        target = self.spec.vector_info(t, ti, tf, self.batch.xpos(), self.tables).get(name) if name in self.spec.info_keys else None
        if target is None:
            raise AttributeError(f"{self.spec.name}: {name!r} is not a settable per-env attribute (per-env: {self.spec.info_keys})")

        def write(col, vals, what):
            if getattr(col, "_base", None) is not ti and getattr(col, "_base", None) is not tf:
                raise AttributeError(f"{self.spec.name}: {what!r} is derived from the task state, not stored in it; it cannot be set")
            col.copy_(t.as_tensor(vals, device=col.device).to(col.dtype).expand_as(col))

        if isinstance(target, dict):
            rows = [values] * self.num_envs if isinstance(values, dict) else list(values)
            if len(rows) != self.num_envs:
                raise ValueError(f"{name}: expected {self.num_envs} dicts, got {len(rows)}")
            for k in rows[0]:
                if k not in target:
                    raise AttributeError(f"{self.spec.name}: {name!r} has no entry {k!r}")
                write(target[k], [r[k] for r in rows], f"{name}.{k}")
        else:
            if not isinstance(values, (int, float)) and len(values) != self.num_envs:
                raise ValueError(f"{name}: expected {self.num_envs} values, got {len(values)}")
            write(target, values, name)
        self.batch.set_task_state(ti, tf)

    def call(self, name: str, *args, **kwargs):
        """``AsyncVectorEnv.call``: per-env results of a method, or the attribute itself when ``name`` is not callable."""
        if name == "render":
            return (None,) * self.num_envs
        return self.get_attr(name)

    @property
    def unwrapped(self):
        return self

    def close_extras(self, **kwargs):
        pass

    def task_state(self, with_xpos: bool = False):
        """Per-env task state (ti, tf[, xpos]); rows of envs whose episode ended in the last step hold the finished episode's
        values (snapshot taken in the kernel before the same-step auto-reset), as the reference's terminal ``info`` does."""
        t = self.torch
        ti, tf = self.batch.get_task_state()
        x = self.batch.xpos() if with_xpos else None
        if bool(self._done.any()):
            fti, ftf, fx = self.batch.final_state()
            d = self._done
            ti = t.where(d[:, None], fti, ti); tf = t.where(d[:, None], ftf, tf)
            if with_xpos:
                x = t.where(d[:, None, None], fx, x)
        return (ti, tf, x) if with_xpos else (ti, tf)

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
