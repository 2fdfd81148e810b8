"""ctypes binding of libchol_b200.so (include/chol_b200.h).  Fails loudly when the CUDA
library is missing: there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CHOL_LIB_PATH") or os.path.join(_HERE, "libchol_b200.so")   # CHOL_LIB_PATH: A/B builds

c_void_p, c_int, c_double, c_ll, c_ull, c_size_t = C.c_void_p, C.c_int, C.c_double, C.c_longlong, C.c_ulonglong, C.c_size_t

# name -> (restype, argtypes); every symbol include/chol_b200.h declares
SIGNATURES = {
    "chol_init": (c_int, [c_int]),
    "chol_finalize": (c_int, []),
    "chol_last_error": (C.c_char_p, []),
    "chol_version": (C.c_char_p, []),
    "chol_launch_count": (c_ull, []),
    "chol_gemm_tasks": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_double, c_void_p]),
    "chol_gemm_tasks_ex": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_double, c_int,
                                   c_void_p]),
    "chol_potrf_tile_workspace": (c_size_t, [c_int]),
    "chol_potrf_tile": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "chol_trsm_tile_workspace": (c_size_t, [c_int]),
    "chol_trsm_tile": (c_int, [c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "chol_trsm_tiles": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "chol_trsm_tiles_push": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "chol_syrk_tile": (c_int, [c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "chol_gemm_tile": (c_int, [c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "chol_potrf_batched": (c_int, [c_int, c_int, c_void_p, c_int, c_ll, c_void_p, c_void_p]),
    "chol_peer_alloc": (c_int, [c_size_t, C.POINTER(c_void_p)]),
    "chol_peer_free": (c_int, [c_void_p]),
    "chol_peer_export": (c_int, [c_void_p, c_void_p]),
    "chol_peer_open": (c_int, [c_void_p, C.POINTER(c_void_p)]),
    "chol_peer_close": (c_int, [c_void_p]),
    "chol_peer_send": (c_int, [c_void_p, c_int, c_void_p]),
    "chol_flag_wait": (c_int, [c_void_p, C.c_uint32, c_void_p]),
    "chol_flag_post": (c_int, [c_void_p, c_int, C.c_uint32, c_void_p]),
    "chol_partition_create": (c_int, [c_int, c_int, c_void_p]),
    "chol_partition_destroy": (c_int, [c_void_p]),
    "chol_plgsy_tile": (c_int, [c_double, c_int, c_int, c_void_p, c_int, c_ll, c_ll, c_ll, c_ll, c_ull, c_void_p]),
    "chol_tile_sumsq": (c_int, [c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "chol_tile_abs_sums": (c_int, [c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "chol_tile_tril": (c_int, [c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "chol_tile_transpose": (c_int, [c_int, c_void_p, c_int, c_ll, c_int, c_void_p]),
    "chol_fp64_peak": (c_int, [c_int, c_int, C.POINTER(c_double), c_void_p]),
}


class Xfer(C.Structure):
    """chol_xfer_t (include/chol_b200.h)."""
    _fields_ = [("dst", c_void_p), ("src", c_void_p), ("tile_bytes", c_ll), ("count", c_int), ("dst_stride", c_int),
                ("src_stride", c_int), ("credit_value", C.c_uint32), ("credit", c_void_p), ("flag", c_void_p),
                ("flag_value", C.c_uint32), ("reserved", C.c_uint32), ("stream", c_void_p)]


class Partition(C.Structure):
    """chol_partition_t (include/chol_b200.h)."""
    _fields_ = [("handle", c_void_p), ("panel_stream", c_void_p), ("rest_stream", c_void_p), ("panel_sms", c_int),
                ("rest_sms", c_int)]


class CholError(RuntimeError):
    """A C-ABI call returned non-zero (argument error < 0, CUDA error > 0)."""

    def __init__(self, fn: str, rc: int, msg: str):
        super().__init__(f"{fn} failed (rc={rc}): {msg}")
        self.fn, self.rc, self.msg = fn, rc, msg


_lib = None


def load() -> C.CDLL:
    """Load libchol_b200.so and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with dense-linear-app_b200/csrc/build.sh "
            "(or __graft_entry__.build()).  There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Call an int-returning entry point and raise CholError on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise CholError(name, rc, lib.chol_last_error().decode())


def version() -> str:
    return load().chol_version().decode()
