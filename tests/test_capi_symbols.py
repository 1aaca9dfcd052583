"""CPU-side: the C-ABI library loads and exports every symbol include/bezgpu.h
declares (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "optimalbeziertrajectorygeneration_b200", "libbezgpu.so")


def _ensure_built():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bezgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bez_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    _ensure_built()
    lib = ctypes.CDLL(LIB)
    syms = declared_symbols()
    assert len(syms) >= 8
    for s in syms:
        assert hasattr(lib, s), "libbezgpu.so does not export %s" % s


def test_binding_covers_header():
    _ensure_built()
    from optimalbeziertrajectorygeneration_b200 import _capi
    assert set(declared_symbols()) <= set(_capi.SIGNATURES)
    assert _capi.lib.bez_version() >= 100


def test_argument_errors_without_gpu():
    """argument validation happens before any CUDA call"""
    _ensure_built()
    from optimalbeziertrajectorygeneration_b200 import _capi
    h = ctypes.c_void_p(0)
    rc = _capi.lib.bez_plan_create(99, 3, 0, 0, None, None, None, ctypes.byref(h))
    assert rc < 0
    assert "NULL" in _capi.last_error() or "degree" in _capi.last_error()


def test_product_path_fails_loudly_without_cuda():
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from optimalbeziertrajectorygeneration_b200 import _capi, optimization
    b = optimization.BezOptimization(numVeh=2, dimension=2, degree=3, initPoints=[(0, 0), (1, 1)],
                                     finalPoints=[(2, 2), (3, 3)])
    with pytest.raises(_capi.BezGpuError):
        b.temporalSeparationConstraints(np.zeros(b.nvar))
