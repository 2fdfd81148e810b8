"""The C-ABI library loads without a GPU and exports every symbol include/chol_b200.h declares;
argument validation happens before any CUDA call.  No compute here."""
import os
import re

import pytest

from dense_linear_app_b200 import _lib

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "chol_b200.h")


def declared_symbols():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(chol_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in chol_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_workspace_queries():
    assert "sm_100a" in _lib.version()
    lib = _lib.load()
    assert lib.chol_potrf_tile_workspace(1024) >= 8 * 128 * 128 * 8
    assert lib.chol_potrf_tile_workspace(0) == 0


def test_argument_errors_are_lapack_style():
    lib = _lib.load()
    assert lib.chol_potrf_tile(-1, None, 1, None, None, 0, None) == -1
    assert lib.chol_potrf_tile(4, None, 4, None, None, 0, None) == -2
    assert lib.chol_gemm_tile(4, None, 4, None, 4, None, 4, None) == -2
    assert b"bad argument" in lib.chol_last_error()
    with pytest.raises(_lib.CholError):
        _lib.call("chol_syrk_tile", 8, None, 8, None, 8, None)


def test_no_cpu_fallback():
    """Without a CUDA device a compute call must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.CholError):
        _lib.call("chol_init", 0)
    from dense_linear_app_b200 import tile_ops
    t = torch.eye(4, dtype=torch.float64)
    with pytest.raises(_lib.CholError):
        tile_ops.potrf_tile(t)
