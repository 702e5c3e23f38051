"""CPU-side checks of the drop-in boundary: libcpm.so builds for sm_100a, loads without a GPU, exports every
symbol include/cpm.h declares, and FAILS LOUDLY (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from carparkingmaps_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(L.SO_PATH):
        L.build()
    return L.lib()


def _declared():
    hdr = open(os.path.join(ROOT, "include", "cpm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cpm_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cpm.h but not exported by libcpm.so"
    assert set(names) == set(L.SIGNATURES), "ctypes signature table out of sync with include/cpm.h"
    assert lib.cpm_version() == 100


def test_header_cites_the_reference(lib):
    hdr = open(os.path.join(ROOT, "include", "cpm.h")).read()
    for cite in ["src/createpdrive.jl:3-38", "src/createpdestin.jl:3-50", "src/initializestates.jl:4-22",
                 "src/solveinitialvalueproblem.jl:4-62", "src/resampling.jl:3-89", "src/averagedrivingtime.jl:3-12",
                 "src/correctparameters.jl:3-23", "src/saveresults.jl:6-28"]:
        assert cite in hdr


def test_correct_parameters_is_pure_host_arithmetic(lib):
    out = np.zeros(2)
    assert lib.cpm_correct_parameters(-0.5, 1.5, 0.1, 0.9, out.ctypes.data) == 0
    assert tuple(out) == (0.0, 1.0)
    assert lib.cpm_correct_parameters(0.95, 0.05, 0.1, 0.9, out.ctypes.data) == 0
    assert tuple(out) == (0.9, 0.1)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is for GPU-less hosts")
    h = C.c_void_p()
    rc = lib.cpm_ctx_create(0, None, C.byref(h))
    assert rc == 2 and not h.value          # CPM_ERR_CUDA
    assert lib.cpm_last_error()
    from carparkingmaps_b200 import Context, CpmError
    with pytest.raises(CpmError):
        Context(0)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "carparkingmaps_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "cpm_oracle" not in src, f
