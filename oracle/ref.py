"""ctypes binding of the fp64 CPU oracle (oracle/mjstep_ref.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  ``RefModel``/``RefData`` mimic the ``mujoco.MjModel``/``MjData`` attribute surface the
reference envs touch (qpos, qvel, ctrl, xpos, xquat, ncon, contact[i].geom1 ...), so the task
restatements in ``oracle/tasks_ref.py`` read like the reference's own env code.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    so = os.path.join(_HERE, "libmjstep_ref.so")
    src = os.path.join(_HERE, "mjstep_ref.c")
    if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.ref_model_create.restype = ctypes.c_void_p
        L.ref_model_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.ref_data_create.restype = ctypes.c_void_p
        L.ref_data_create.argtypes = [ctypes.c_void_p]
        for f in ("ref_model_destroy", "ref_data_destroy"):
            getattr(L, f).argtypes = [ctypes.c_void_p]
        for f in ("ref_reset_data", "ref_forward", "ref_step"):
            getattr(L, f).argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ref_step_n.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.ref_field.restype = ctypes.POINTER(ctypes.c_double)
        L.ref_field.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
        for f in ("ref_ncon", "ref_nefc", "ref_solver_iter", "ref_nwarn"):
            getattr(L, f).argtypes = [ctypes.c_void_p]; getattr(L, f).restype = ctypes.c_int
        L.ref_set_flags.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.ref_set_warmstart_mode.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.ref_contact.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _LIB = L
    return _LIB


class RefModel:
    """Oracle-side model: the packed tables plus the name lookups the reference does with mj_name2id."""

    def __init__(self, tables, ints: np.ndarray, flts: np.ndarray):
        self.tables = tables
        self._ints = np.ascontiguousarray(ints, np.int32)
        self._flts = np.ascontiguousarray(flts, np.float64)
        self.ptr = lib().ref_model_create(self._ints.ctypes.data, self._ints.size, self._flts.ctypes.data, self._flts.size)
        if not self.ptr:
            raise ValueError("packed model rejected by the oracle (layout mismatch)")
        for k in ("nq", "nv", "nu", "nbody", "njnt", "ngeom"):
            setattr(self, k, int(getattr(tables, k)))

    def __del__(self):
        if getattr(self, "ptr", None) and _LIB is not None:
            _LIB.ref_model_destroy(self.ptr); self.ptr = None


class _Contact:
    __slots__ = ("geom1", "geom2", "dim", "efc_address", "dist", "pos", "frame", "friction")


class RefData:
    _SHAPES = {"xpos": 3, "xquat": 4, "xmat": 9, "xipos": 3, "ximat": 9, "geom_xpos": 3, "geom_xmat": 9,
               "site_xpos": 3, "subtree_com": 3, "cvel": 6, "cdof": 6, "cinert": 10, "xfrc_applied": 6}
    _STATIC = ("qpos", "qvel", "ctrl", "qfrc_applied", "xfrc_applied", "qacc", "qacc_warmstart", "xpos", "xquat",
               "xmat", "xipos", "ximat", "geom_xpos", "geom_xmat", "site_xpos", "subtree_com", "cvel", "cdof",
               "cinert", "qfrc_bias", "qfrc_passive", "qfrc_actuator", "qfrc_smooth", "qacc_smooth",
               "qfrc_constraint")

    def __init__(self, model: RefModel):
        self.model = model
        self.ptr = lib().ref_data_create(model.ptr)
        self._views = {}
        for name in self._STATIC:
            self._views[name] = self._view(name)

    def _view(self, name):
        n = ctypes.c_int(0)
        p = lib().ref_field(self.model.ptr, self.ptr, name.encode(), ctypes.byref(n))
        if not p or n.value == 0:
            return np.zeros(0)
        a = np.ctypeslib.as_array(p, shape=(n.value,))
        w = self._SHAPES.get(name)
        return a.reshape(-1, w) if w else a

    def __getattr__(self, name):
        views = self.__dict__.get("_views", {})
        if name in views:
            return views[name]
        if name in ("efc_force", "efc_pos", "efc_R", "efc_D", "efc_aref", "efc_b"):
            return self._view(name).copy()
        if name == "efc_J":
            return self._view(name).reshape(self.nefc, self.model.nv).copy()
        if name == "efc_AR":
            return self._view(name).reshape(self.nefc, self.nefc).copy()
        if name == "M":
            return self._view("M").reshape(self.model.nv, self.model.nv).copy()
        if name == "time":
            return float(self._view("time")[0])
        raise AttributeError(name)

    def set_warmstart_once_per_step(self, on: bool):
        """False (default): qacc_warmstart saved by every mj_fwdConstraint (MuJoCo 3.x); True: once per mj_step."""
        lib().ref_set_warmstart_mode(self.ptr, int(bool(on)))

    @property
    def ncon(self): return lib().ref_ncon(self.ptr)
    @property
    def nefc(self): return lib().ref_nefc(self.ptr)
    @property
    def solver_iter(self): return lib().ref_solver_iter(self.ptr)
    @property
    def nwarn(self): return lib().ref_nwarn(self.ptr)

    @property
    def contact(self):
        out = []
        gi = (ctypes.c_int * 4)(); buf = (ctypes.c_double * 18)()
        for k in range(self.ncon):
            lib().ref_contact(self.ptr, k, gi, buf)
            c = _Contact()
            c.geom1, c.geom2, c.dim, c.efc_address = gi[0], gi[1], gi[2], gi[3]
            b = np.array(buf[:])
            c.dist = b[0]; c.pos = b[1:4]; c.frame = b[4:13]; c.friction = b[13:18]
            out.append(c)
        return out

    def __del__(self):
        if getattr(self, "ptr", None) and _LIB is not None:
            _LIB.ref_data_destroy(self.ptr); self.ptr = None


def mj_step(model: RefModel, data: RefData, n: int = 1):
    lib().ref_step_n(model.ptr, data.ptr, n)


def mj_forward(model: RefModel, data: RefData):
    lib().ref_forward(model.ptr, data.ptr)


def mj_resetData(model: RefModel, data: RefData):
    lib().ref_reset_data(model.ptr, data.ptr)


def load_model(tables) -> RefModel:
    """Build the oracle model from ModelTables using the same packed buffers the C-ABI takes."""
    from mujoco_gymnasium_environments_b200.model_pack import pack_model
    ints, flts = pack_model(tables)
    return RefModel(tables, ints, flts)
