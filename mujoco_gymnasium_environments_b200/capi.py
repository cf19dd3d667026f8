"""ctypes binding of libb2env.so (the C-ABI in include/b2env.h) plus a thin torch-tensor wrapper.

The product path fails loudly when the CUDA library is missing or no GPU is present: there is no CPU fallback
and nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("B2_LIB_PATH", os.path.join(_HERE, "libb2env.so"))      # override: A/B builds in profiling runs
_LIB = None

SYMBOLS = ["b2_model_create", "b2_model_destroy", "b2_batch_create", "b2_batch_destroy", "b2_dims", "b2_reset",
           "b2_step", "b2_step_host", "b2_physics_step", "b2_forward", "b2_get_state", "b2_set_state",
           "b2_get_task_state", "b2_set_task_state", "b2_get_contacts", "b2_get_xpos", "b2_debug_forward", "b2_stats",
           "b2_launch_count", "b2_last_error", "b2_rollout", "b2_host_buffers", "b2_reseed", "b2_caps", "b2_get_final_state"]

TASK_NONE, TASK_QUADRUPED_PARKOUR, TASK_HUMANOID_DANCING, TASK_HUMANOID_SOCCER, TASK_BIPEDAL_RESCUE, TASK_HUMANOID_CONSTRUCTION, TASK_HUMANOID_MARTIAL_ARTS, TASK_ROBOTIC_ARM_ASSEMBLY = 0, 1, 2, 3, 4, 5, 6, 7


class B2TaskDesc(ctypes.Structure):
    _fields_ = [("task", ctypes.c_int), ("ids", ctypes.c_int * 16), ("act_lo", ctypes.c_float * 40),
                ("act_hi", ctypes.c_float * 40), ("aux_i", ctypes.c_int * 64), ("aux_f", ctypes.c_float * 32)]


class B2BatchOpts(ctypes.Structure):
    _fields_ = [("envs_per_block", ctypes.c_int), ("arena_floats", ctypes.c_int), ("con_cap", ctypes.c_int),
                ("row_cap", ctypes.c_int), ("warps_per_env", ctypes.c_int), ("disable_wide", ctypes.c_int),
                ("warmstart_once_per_step", ctypes.c_int), ("fifo_queue", ctypes.c_int)]


class B2Error(RuntimeError):
    pass


def build(force: bool = False) -> str:
    """Compile csrc/ into libb2env.so for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh"))]
    srcs += [os.path.join(_HERE, "..", "include", f) for f in ("b2env.h", "b2_device_layout.h")]
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", src_dir, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            raise B2Error(f"{_SO} is missing: run __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback.")
        L = ctypes.CDLL(_SO)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        L.b2_model_create.argtypes = [vp, ci, vp, ci, ci, ctypes.POINTER(vp)]
        L.b2_model_destroy.argtypes = [vp]; L.b2_model_destroy.restype = None
        L.b2_batch_create.argtypes = [vp, ctypes.POINTER(B2TaskDesc), ci, ctypes.c_uint64, ci, ctypes.POINTER(B2BatchOpts), ctypes.POINTER(vp)]
        L.b2_batch_destroy.argtypes = [vp]; L.b2_batch_destroy.restype = None
        L.b2_dims.argtypes = [vp, vp]
        L.b2_reset.argtypes = [vp, vp, vp, vp, vp]
        L.b2_step.argtypes = [vp] * 8
        L.b2_step_host.argtypes = [vp] * 6
        L.b2_physics_step.argtypes = [vp, ci, vp]
        L.b2_forward.argtypes = [vp, vp]
        L.b2_get_state.argtypes = [vp] * 7; L.b2_set_state.argtypes = [vp] * 7
        L.b2_get_task_state.argtypes = [vp] * 4; L.b2_set_task_state.argtypes = [vp] * 4
        L.b2_get_contacts.argtypes = [vp, vp, vp, vp, ci, vp]
        L.b2_get_xpos.argtypes = [vp, vp, vp]
        L.b2_debug_forward.argtypes = [vp, vp, ci, vp]
        L.b2_stats.argtypes = [vp, vp, vp]
        L.b2_rollout.argtypes = [vp, ci] + [vp] * 7
        L.b2_host_buffers.argtypes = [vp] + [ctypes.POINTER(vp)] * 5
        L.b2_reseed.argtypes = [vp, ctypes.c_uint64, vp]
        L.b2_caps.argtypes = [vp, vp]
        L.b2_get_final_state.argtypes = [vp] * 5
        L.b2_launch_count.restype = ctypes.c_ulonglong
        L.b2_last_error.restype = ctypes.c_char_p
        _LIB = L
    return _LIB


def _ck(rc: int):
    if rc != 0:
        raise B2Error(f"libb2env error {rc}: {lib().b2_last_error().decode()}")


def launch_count() -> int:
    return int(lib().b2_launch_count())


class DeviceModel:
    """``mujoco.MjModel`` stand-in living on one GPU."""

    def __init__(self, tables, device: int = 0):
        from .device_pack import pack_device_model
        self.tables = tables
        ints, flts = pack_device_model(tables)
        self._ints = np.ascontiguousarray(ints, np.int32); self._flts = np.ascontiguousarray(flts, np.float64)
        self.device = device
        h = ctypes.c_void_p()
        _ck(lib().b2_model_create(self._ints.ctypes.data, self._ints.size, self._flts.ctypes.data, self._flts.size,
                                  device, ctypes.byref(h)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            lib().b2_model_destroy(self.handle); self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ptr(t):
    return ctypes.c_void_p(0) if t is None else ctypes.c_void_p(t.data_ptr())


class Batch:
    """``mujoco.MjData`` x n_envs stand-in; all tensors are torch CUDA tensors on the model's device."""

    def __init__(self, model: DeviceModel, task: Optional[B2TaskDesc], n_envs: int, seed: int = 0, env_offset: int = 0,
                 envs_per_block: int = 0, arena_floats: int = 0, con_cap: int = 0, row_cap: int = 0,
                 warps_per_env: int = 0, disable_wide: bool = False, warmstart_once_per_step: bool = False, fifo_queue: bool = False):
        import torch
        self.torch = torch
        self.model = model
        self.device = torch.device("cuda", model.device)
        h = ctypes.c_void_p()
        opts = B2BatchOpts()
        opts.envs_per_block = int(os.environ.get("B2_EPB", envs_per_block)); opts.arena_floats = int(os.environ.get("B2_ARENA", arena_floats))
        opts.con_cap = int(os.environ.get("B2_CON_CAP", con_cap)); opts.row_cap = int(os.environ.get("B2_ROW_CAP", row_cap))
        opts.warps_per_env = int(os.environ.get("B2_WPE", warps_per_env))
        opts.disable_wide = int(os.environ.get("B2_NO_WIDE", int(disable_wide)))
        opts.warmstart_once_per_step = int(os.environ.get("B2_WARM_ONCE", int(warmstart_once_per_step)))
        opts.fifo_queue = int(os.environ.get("B2_FIFO", int(fifo_queue)))
        _ck(lib().b2_batch_create(model.handle, ctypes.byref(task) if task is not None else None, n_envs,
                                  ctypes.c_uint64(seed & (2**64 - 1)), env_offset, ctypes.byref(opts), ctypes.byref(h)))
        self.handle = h
        d = (ctypes.c_int * 16)()
        _ck(lib().b2_dims(self.handle, d))
        (self.nq, self.nv, self.nu, self.nbody, self.obs_dim, self.act_dim, self.n_envs, self.nti, self.ntf, self.con_cap,
         self.smem_bytes, self.envs_per_block, self.row_cap, self.nM, self.arena_floats, self.ws_bytes) = [int(x) for x in d[:16]]
        _ck(lib().b2_caps(self.handle, d))
        (self.ninj, self.wide_con_cap, self.wide_row_cap, self.wide_arena_floats, self.wide_kib_per_env, self.raw_cap, self.act_cap,
         self.wide_enabled, self.episode_slot, self.warm_once) = [int(x) for x in d[:10]]
        self._host = None

    # ---- argument checks: the kernel indexes these buffers by env without bounds checks (a wrong shape would read out of bounds)
    def _check(self, name, t, shape, dtype=None):
        if t is None:
            return
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
        if t.device != self.device:
            raise ValueError(f"{name}: expected a tensor on {self.device}, got {t.device}")
        if dtype is not None and t.dtype != dtype:
            raise ValueError(f"{name}: expected dtype {dtype}, got {t.dtype}")
        if not t.is_contiguous():
            raise ValueError(f"{name}: must be contiguous")

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _new(self, shape, dtype=None):
        return self.torch.empty(shape, dtype=dtype or self.torch.float32, device=self.device)

    # ---- hot path
    def reset(self, obs, mask=None, inject=None):
        t = self.torch
        self._check("obs", obs, (self.n_envs, self.obs_dim), t.float32)
        if mask is not None:
            self._check("reset_mask", mask, (self.n_envs,), t.uint8)
        if inject is not None:
            self._check("inject", inject, (self.n_envs, self.ninj), t.float32)
        _ck(lib().b2_reset(self.handle, _ptr(mask), _ptr(inject), _ptr(obs), self._stream()))

    def reseed(self, seed: int):
        """``Env.reset(seed=...)``: new RNG seed, episode counters restart (same seed -> same initial states)."""
        _ck(lib().b2_reseed(self.handle, ctypes.c_uint64(int(seed) & (2**64 - 1)), self._stream()))

    def step(self, act, obs, rew, term, trunc, final_obs=None):
        t = self.torch
        self._check("actions", act, (self.n_envs, self.act_dim), t.float32); self._check("obs", obs, (self.n_envs, self.obs_dim), t.float32)
        self._check("reward", rew, (self.n_envs,), t.float32); self._check("terminated", term, (self.n_envs,), t.uint8)
        self._check("truncated", trunc, (self.n_envs,), t.uint8); self._check("final_obs", final_obs, (self.n_envs, self.obs_dim), t.float32)
        _ck(lib().b2_step(self.handle, _ptr(act), _ptr(obs), _ptr(rew), _ptr(term), _ptr(trunc), _ptr(final_obs), self._stream()))

    def rollout(self, act, obs, rew, term, trunc, final_obs=None):
        """T env.steps with the host out of the loop (one CUDA graph): act [T, N, A] -> obs [T, N, D], rew/term/trunc [T, N]."""
        t = self.torch; T = int(act.shape[0])
        self._check("actions", act, (T, self.n_envs, self.act_dim), t.float32); self._check("obs", obs, (T, self.n_envs, self.obs_dim), t.float32)
        self._check("reward", rew, (T, self.n_envs), t.float32); self._check("terminated", term, (T, self.n_envs), t.uint8)
        self._check("truncated", trunc, (T, self.n_envs), t.uint8); self._check("final_obs", final_obs, (T, self.n_envs, self.obs_dim), t.float32)
        _ck(lib().b2_rollout(self.handle, T, _ptr(act), _ptr(obs), _ptr(rew), _ptr(term), _ptr(trunc), _ptr(final_obs), self._stream()))

    def step_host(self, act: np.ndarray, obs: np.ndarray, rew: np.ndarray, term: np.ndarray, trunc: np.ndarray):
        for name, a, shape, dt in (("actions", act, (self.n_envs, self.act_dim), np.float32), ("obs", obs, (self.n_envs, self.obs_dim), np.float32),
                                   ("reward", rew, (self.n_envs,), np.float32), ("terminated", term, (self.n_envs,), np.uint8),
                                   ("truncated", trunc, (self.n_envs,), np.uint8)):
            if a.shape != shape or a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                raise ValueError(f"{name}: expected a C-contiguous {np.dtype(dt).name} array of shape {shape}, got {a.dtype} {a.shape}")
        _ck(lib().b2_step_host(self.handle, act.ctypes.data, obs.ctypes.data, rew.ctypes.data, term.ctypes.data, trunc.ctypes.data))

    def host_buffers(self):
        """numpy views of the library's pinned staging buffers (act, obs, rew, term, trunc): pass them to step_host for zero-copy."""
        if self._host is None:
            p = [ctypes.c_void_p() for _ in range(5)]
            _ck(lib().b2_host_buffers(self.handle, *[ctypes.byref(x) for x in p]))
            def view(ptr, shape, ctype, dt):
                n = int(np.prod(shape)); buf = (ctype * n).from_address(ptr.value)
                return np.frombuffer(buf, dtype=dt).reshape(shape)
            N = self.n_envs
            self._host = (view(p[0], (N, max(self.act_dim, 1)), ctypes.c_float, np.float32)[:, :self.act_dim] if self.act_dim else None,
                          view(p[1], (N, max(self.obs_dim, 1)), ctypes.c_float, np.float32), view(p[2], (N,), ctypes.c_float, np.float32),
                          view(p[3], (N,), ctypes.c_uint8, np.uint8), view(p[4], (N,), ctypes.c_uint8, np.uint8))
        return self._host

    def final_state(self):
        """Task state / xpos of the episode that ended in each env's last terminal step (before the same-step auto-reset)."""
        ti = self._new((self.n_envs, max(self.nti, 1)), self.torch.int32); tf = self._new((self.n_envs, max(self.ntf, 1)))
        x = self._new((self.n_envs, self.nbody, 3))
        _ck(lib().b2_get_final_state(self.handle, _ptr(ti), _ptr(tf), _ptr(x), self._stream()))
        return ti, tf, x

    def physics_step(self, nsub: int = 1):
        _ck(lib().b2_physics_step(self.handle, nsub, self._stream()))

    def forward(self):
        _ck(lib().b2_forward(self.handle, self._stream()))

    # ---- state access
    def get_state(self):
        q, v, c, w, t = (self._new((self.n_envs, self.nq)), self._new((self.n_envs, self.nv)),
                         self._new((self.n_envs, max(self.nu, 1))), self._new((self.n_envs, self.nv)), self._new((self.n_envs,)))
        _ck(lib().b2_get_state(self.handle, _ptr(q), _ptr(v), _ptr(c) if self.nu else None, _ptr(w), _ptr(t), self._stream()))
        return dict(qpos=q, qvel=v, ctrl=c[:, :self.nu], qacc_warmstart=w, time=t)

    def set_state(self, qpos=None, qvel=None, ctrl=None, qacc_warmstart=None, time=None):
        f = lambda x: None if x is None else x.to(self.device, self.torch.float32).contiguous()
        q, v, c, w, t = f(qpos), f(qvel), f(ctrl), f(qacc_warmstart), f(time)
        _ck(lib().b2_set_state(self.handle, _ptr(q), _ptr(v), _ptr(c), _ptr(w), _ptr(t), self._stream()))
        self.torch.cuda.current_stream(self.device).synchronize()

    def get_task_state(self):
        ti = self._new((self.n_envs, max(self.nti, 1)), self.torch.int32); tf = self._new((self.n_envs, max(self.ntf, 1)))
        _ck(lib().b2_get_task_state(self.handle, _ptr(ti), _ptr(tf), self._stream()))
        return ti, tf

    def set_task_state(self, ti=None, tf=None):
        ti = None if ti is None else ti.to(self.device, self.torch.int32).contiguous()
        tf = None if tf is None else tf.to(self.device, self.torch.float32).contiguous()
        _ck(lib().b2_set_task_state(self.handle, _ptr(ti), _ptr(tf), self._stream()))
        self.torch.cuda.current_stream(self.device).synchronize()

    def contacts(self, cap: Optional[int] = None):
        cap = cap or self.con_cap
        ncon = self._new((self.n_envs,), self.torch.int32); geom = self._new((self.n_envs, cap, 2), self.torch.int32)
        dist = self._new((self.n_envs, cap))
        geom.fill_(-1); dist.zero_()
        _ck(lib().b2_get_contacts(self.handle, _ptr(ncon), _ptr(geom), _ptr(dist), cap, self._stream()))
        return ncon, geom, dist

    def xpos(self):
        x = self._new((self.n_envs, self.nbody, 3))
        _ck(lib().b2_get_xpos(self.handle, _ptr(x), self._stream()))
        return x

    def debug_forward(self):
        n = 4 * self.nv + self.nM + 4 + 4 * self.row_cap
        out = self._new((self.n_envs, n)); out.zero_()
        _ck(lib().b2_debug_forward(self.handle, _ptr(out), n, self._stream()))
        nv, nM, rc = self.nv, self.nM, self.row_cap
        o = out
        k = 4 * nv + nM
        return dict(qfrc_smooth=o[:, 0:nv], qacc_smooth=o[:, nv:2 * nv], qfrc_constraint=o[:, 2 * nv:3 * nv],
                    qacc=o[:, 3 * nv:4 * nv], M=o[:, 4 * nv:k], ncon=o[:, k].int(), nefc=o[:, k + 1].int(),
                    solver_iter=o[:, k + 2].int(), efc_force=o[:, k + 4:k + 4 + rc], efc_b=o[:, k + 4 + rc:k + 4 + 2 * rc],
                    efc_R=o[:, k + 4 + 2 * rc:k + 4 + 3 * rc], efc_pos=o[:, k + 4 + 3 * rc:k + 4 + 4 * rc])

    def stats(self):
        out = self._new((16,), self.torch.float64)
        _ck(lib().b2_stats(self.handle, _ptr(out), self._stream()))
        return out

    def close(self):
        if getattr(self, "handle", None):
            lib().b2_batch_destroy(self.handle); self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
