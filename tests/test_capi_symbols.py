"""The C-ABI library loads on a CPU-only box and exports every symbol include/b2env.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from mujoco_gymnasium_environments_b200 import capi

ROOT = os.path.join(os.path.dirname(__file__), "..")


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b2env.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    so = capi.build()
    assert os.path.exists(so)
    L = ctypes.CDLL(so)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/b2env.h but not exported"
    assert sorted(capi.SYMBOLS) == syms


def test_struct_layouts_match_header():
    assert ctypes.sizeof(capi.B2TaskDesc) == 4 + 16 * 4 + 40 * 4 + 40 * 4 + 64 * 4 + 32 * 4
    assert ctypes.sizeof(capi.B2BatchOpts) == 8 * 4


def test_errors_do_not_cross_the_abi():
    L = capi.lib()
    h = ctypes.c_void_p()
    rc = L.b2_model_create(None, 0, None, 0, 0, ctypes.byref(h))
    assert rc == -1 and b"null" in L.b2_last_error()
    import numpy as np
    bad = np.zeros(8, np.int32); f = np.zeros(4)
    rc = L.b2_model_create(bad.ctypes.data, 8, f.ctypes.data, 4, 0, ctypes.byref(h))
    assert rc == -3


def test_no_cpu_fallback_in_product_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mujoco_gymnasium_environments_b200.vector_env import B200VectorEnv
    with pytest.raises(capi.B2Error):
        B200VectorEnv("quadruped_parkour", 4)
    # the product package never imports the oracle
    pkg = os.path.join(ROOT, "mujoco_gymnasium_environments_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
