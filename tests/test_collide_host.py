"""Differential test of the product's fp32 narrow phase (csrc/b2_collide.cuh, compiled for the host with the CUDA
qualifiers defined away) against the fp64 oracle (oracle/mjstep_ref.c::ref_collide_raw), pair type by pair type, on
random and axis-aligned configurations.  Runs without a GPU; the same primitives run on the device in the -m gpu tests.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from mujoco_gymnasium_environments_b200.mjcf import quat_to_mat
from oracle import ref

HERE = os.path.dirname(os.path.abspath(__file__))
FP = ctypes.POINTER(ctypes.c_float); DP = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("collide") / "libcollide_host.so")
    subprocess.check_call(["g++", "-O1", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so,
                           os.path.join(HERE, "host", "collide_host.cpp")])
    H = ctypes.CDLL(so)
    H.h_collide_pair.argtypes = [ctypes.c_int, ctypes.c_int, FP, FP, FP, FP, FP, FP, ctypes.c_float, FP]
    R = ref.lib()
    R.ref_collide_raw.argtypes = [ctypes.c_int, ctypes.c_int, DP, DP, DP, DP, DP, DP, ctypes.c_double, DP]
    return H, R


def both(H, R, t1, t2, p1, m1, s1, p2, m2, s2, margin):
    a = [np.ascontiguousarray(x, np.float32) for x in (p1, m1, s1, p2, m2, s2)]
    b = [np.ascontiguousarray(x, np.float64) for x in a]        # the oracle sees the fp32-rounded inputs
    o32 = np.zeros(80, np.float32); o64 = np.zeros(80)
    n32 = H.h_collide_pair(t1, t2, *[x.ctypes.data_as(FP) for x in a], margin, o32.ctypes.data_as(FP))
    n64 = R.ref_collide_raw(t1, t2, *[x.ctypes.data_as(DP) for x in b], float(np.float32(margin)), o64.ctypes.data_as(DP))
    return n32, o32.reshape(8, 10), n64, o64.reshape(8, 10)


SIZES = {0: lambda r: [0, 0, 0], 2: lambda r: [r.uniform(.03, .2), 0, 0], 3: lambda r: [r.uniform(.03, .1), r.uniform(.05, .3), 0],
         5: lambda r: [r.uniform(.5, 3), r.uniform(.05, .2), 0], 6: lambda r: list(r.uniform(.03, .25, 3))}
PAIRS = [(0, 2), (0, 3), (0, 5), (0, 6), (2, 2), (2, 3), (2, 6), (2, 5), (3, 3), (3, 5), (5, 5), (5, 6), (3, 6), (6, 6)]
# cylinders come in two shapes: flat discs (dancing floor, rescue wheels) and the martial-arts dummies (r 0.2, half height 0.5)
TALL = lambda r: [r.uniform(.1, .3), r.uniform(.3, .6), 0]


@pytest.mark.parametrize("pair", PAIRS)
def test_fp32_narrow_phase_matches_oracle(libs, pair):
    H, R = libs
    t1, t2 = pair
    rng = np.random.default_rng(100 * t1 + t2)

    def rq():
        q = rng.normal(size=4); return q / np.linalg.norm(q)

    ncontact = 0
    for it in range(3000):
        s1 = SIZES[t1](rng); s2 = SIZES[t2](rng)
        tall = it % 2 == 1
        if tall and t1 == 5: s1 = TALL(rng)
        if tall and t2 == 5: s2 = TALL(rng)
        m1 = quat_to_mat(rq()).ravel(); m2 = quat_to_mat(rq()).ravel()
        if it % 4 == 0 and t1 != 0:      # axis-aligned / parallel configurations (plateaus, ties)
            m1 = np.eye(3).ravel()
            if it % 8 == 0:
                m2 = np.eye(3).ravel()
        scale = (0.4 if tall else 1.5) if t1 == 5 else 0.3 if t1 == 0 else (0.4 if (tall and t2 == 5) else 0.25)
        p2 = rng.normal(size=3) * scale
        n32, o32, n64, o64 = both(H, R, t1, t2, np.zeros(3), m1, s1, p2, m2, s2, 0.01)
        assert n32 == n64, (pair, it, n32, n64)          # contact counts are bit-exact
        if n64 > 0:
            ncontact += 1
            assert np.abs(o32[:n64, 0] - o64[:n64, 0]).max() < 2e-6, (pair, it)          # dist
            assert np.abs(o32[:n64, 1:7] - o64[:n64, 1:7]).max() < 5e-4, (pair, it)       # pos, normal
    assert ncontact > 100


@pytest.mark.parametrize("pair", [(2, 3), (3, 3), (3, 6), (6, 6), (3, 5), (5, 5), (5, 6)])
def test_convex_path_matches_oracle_on_any_pair(libs, pair):
    """The kernel's MPR (csrc/b2_mpr.cuh, fp64 inside an fp32 interface) against the oracle's on the same fp32-rounded
    inputs, world coordinates of order 1 m, penetrations from grazing to deep: same contacts, results to fp32 output rounding."""
    H, R = libs
    H.h_mpr_pair.argtypes = H.h_collide_pair.argtypes
    R.ref_mpr_raw.argtypes = R.ref_collide_raw.argtypes
    t1, t2 = pair
    rng = np.random.default_rng(7 + 100 * t1 + t2)

    def rq():
        q = rng.normal(size=4); return q / np.linalg.norm(q)

    sizes = dict(SIZES); sizes[5] = TALL
    n = 0
    for it in range(3000):
        s1 = sizes[t1](rng); s2 = sizes[t2](rng)
        m1 = quat_to_mat(rq()).ravel() if it % 4 else np.eye(3).ravel(); m2 = quat_to_mat(rq()).ravel() if it % 8 else np.eye(3).ravel()
        p1 = rng.normal(size=3); p2 = p1 + rng.normal(size=3) * 0.3
        a = [np.ascontiguousarray(x, np.float32) for x in (p1, m1, s1, p2, m2, s2)]
        b = [np.ascontiguousarray(x, np.float64) for x in a]
        o32 = np.zeros(10, np.float32); o64 = np.zeros(10)
        n32 = H.h_mpr_pair(t1, t2, *[x.ctypes.data_as(FP) for x in a], 0.01, o32.ctypes.data_as(FP))
        n64 = R.ref_mpr_raw(t1, t2, *[x.ctypes.data_as(DP) for x in b], float(np.float32(0.01)), o64.ctypes.data_as(DP))
        assert n32 == n64, (pair, it)
        if n64:
            n += 1
            assert abs(o32[0] - o64[0]) < 1e-7 and np.abs(o32[1:4] - o64[1:4]).max() < 5e-7 and np.abs(o32[4:7] - o64[4:7]).max() < 2e-7, (pair, it)
    assert n > 250
