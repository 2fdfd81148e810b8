"""CPU oracle for the tile Cholesky path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product (``dense-linear-app_b200``) never does.

Two layers, both restating the same reference call sites (see chol_oracle.c header for the
file:line map):
  * ``C``      — plain-C loops (oracle/chol_oracle.c -> liboracle_chol.so), the restatement proper;
  * ``blas_*`` — the same four tile ops through scipy's bundled OpenBLAS (the library family the
                 reference itself calls: OpenBLAS dpotrf/dtrsm/dsyrk/dgemm, W2:238/323/416/511),
                 used for larger sizes and as the CPU baseline ("port").

Parity is pinned against the reference's own CPU program compiled here (oracle/_ref, see
pin_against_ref.py and tests/golden/ref_lapacke_dpotrf.json); the reference has no golden
vectors of its own.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_chol.so")
REF_BIN = os.path.join(_HERE, "_ref", "lapacke_dpotrf_ref")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="F_CONTIGUOUS")
_lib = None


def build(force: bool = False) -> None:
    """Compile liboracle_chol.so (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "chol_oracle.c")
    ):
        subprocess.run(["make", "-C", _HERE, "liboracle_chol.so"], check=True, capture_output=True)
    if os.path.exists("/root/reference") and (force or not os.path.exists(REF_BIN)):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.oracle_potrf_tile.restype = C.c_int
        L.oracle_potrf_tile.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.oracle_trsm_tile.restype = None
        L.oracle_trsm_tile.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.oracle_syrk_tile.restype = None
        L.oracle_syrk_tile.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.oracle_gemm_tile.restype = None
        L.oracle_gemm_tile.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.oracle_potrf_tiled.restype = C.c_int
        L.oracle_potrf_tiled.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.oracle_plgsy.restype = None
        L.oracle_plgsy.argtypes = [C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong,
                                   C.c_longlong, C.c_longlong, C.c_ulonglong]
        L.oracle_lp_matrix.restype = None
        L.oracle_lp_matrix.argtypes = [C.c_int, C.c_void_p]
        L.oracle_backward_error.restype = None
        L.oracle_backward_error.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _f(a: np.ndarray) -> np.ndarray:
    assert a.dtype == np.float64 and a.flags.f_contiguous, "oracle works on column-major float64"
    return a


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


# ---- plain-C tile ops (in place, column-major) --------------------------------------------------

def potrf_tile(A: np.ndarray) -> int:
    """A <- chol_lower(A) (W2:238).  Returns LAPACK info."""
    _f(A)
    return lib().oracle_potrf_tile(A.shape[0], _ptr(A), A.strides[1] // 8)


def trsm_tile(L: np.ndarray, A: np.ndarray) -> None:
    """A <- A L^{-T} (W2:323)."""
    _f(L), _f(A)
    lib().oracle_trsm_tile(A.shape[0], A.shape[1], _ptr(L), L.strides[1] // 8, _ptr(A), A.strides[1] // 8)


def syrk_tile(A: np.ndarray, Cm: np.ndarray) -> None:
    """C <- C - A A^T, lower triangle (W2:416)."""
    _f(A), _f(Cm)
    lib().oracle_syrk_tile(Cm.shape[0], A.shape[1], _ptr(A), A.strides[1] // 8, _ptr(Cm), Cm.strides[1] // 8)


def gemm_tile(Ai: np.ndarray, Aj: np.ndarray, Cm: np.ndarray) -> None:
    """C <- C - Ai Aj^T (W2:511)."""
    _f(Ai), _f(Aj), _f(Cm)
    lib().oracle_gemm_tile(Cm.shape[0], Cm.shape[1], Ai.shape[1], _ptr(Ai), Ai.strides[1] // 8, _ptr(Aj),
                           Aj.strides[1] // 8, _ptr(Cm), Cm.strides[1] // 8)


def potrf_tiled(tiles: dict, nt: int, b: int) -> int:
    """Tile DAG in the client's wave order (C1:278-333) on a dict {(i,j): b x b F-array}, i>=j."""
    arr = (C.c_void_p * (nt * (nt + 1) // 2))()
    for i in range(nt):
        for j in range(i + 1):
            t = _f(tiles[(i, j)])
            assert t.shape == (b, b) and t.strides[1] == 8 * b
            arr[i * (i + 1) // 2 + j] = _ptr(t)
    return lib().oracle_potrf_tiled(nt, b, arr)


def plgsy(bump: float, N: int, seed: int, row0: int = 0, col0: int = 0, mb: int | None = None,
          nb: int | None = None, bigM: int | None = None) -> np.ndarray:
    """dplgsy-like SPD generator (V6:46).  Whole matrix by default, or the (mb x nb) tile at (row0,col0)."""
    mb = N if mb is None else mb
    nb = N if nb is None else nb
    A = np.empty((mb, nb), dtype=np.float64, order="F")
    lib().oracle_plgsy(float(bump), mb, nb, _ptr(A), mb, N if bigM is None else bigM, row0, col0, N, seed)
    return A


def plgsy_numpy(bump: float, N: int, seed: int) -> np.ndarray:
    """Independent numpy statement of the same generator (uint64 wrap-around arithmetic)."""
    A_K, C_K = np.uint64(6364136223846793005), np.uint64(1)
    i, j = np.meshgrid(np.arange(N, dtype=np.uint64), np.arange(N, dtype=np.uint64), indexing="ij")
    hi, lo = np.maximum(i, j), np.minimum(i, j)
    n = hi + lo * np.uint64(N)
    ran = np.full((N, N), seed, dtype=np.uint64)
    a_k, c_k = A_K, C_K
    with np.errstate(over="ignore"):
        for _ in range(64):
            odd = (n & np.uint64(1)).astype(bool)
            ran = np.where(odd, a_k * ran + c_k, ran)
            c_k = c_k * (a_k + np.uint64(1))
            a_k = a_k * a_k
            n = n >> np.uint64(1)
    A = 0.5 - ran.astype(np.float64) * 5.4210108624275222e-20
    A[np.arange(N), np.arange(N)] += bump
    return np.asfortranarray(A)


def lp_matrix(N: int) -> np.ndarray:
    """The matrix of the reference's lapacke_dpotrf.c (LP:35-45), as the N x N array it stores
    row-major; symmetric, so the same numbers read column-major."""
    A = np.empty((N, N), dtype=np.float64, order="F")
    lib().oracle_lp_matrix(N, _ptr(A))
    return A


def backward_error(A: np.ndarray, L: np.ndarray) -> float:
    """||A - L L^T||_F / ||A||_F with A symmetric (lower triangle read), L lower (upper ignored)."""
    _f(A), _f(L)
    out = np.zeros(2)
    lib().oracle_backward_error(A.shape[0], _ptr(A), A.strides[1] // 8, _ptr(L), L.strides[1] // 8, _ptr(out))
    return float(out[0] / out[1])


def backward_error_blas(A: np.ndarray, L: np.ndarray) -> float:
    """Same quantity through numpy matmul (for sizes where the plain loop is slow)."""
    Lt = np.tril(L)
    R = Lt @ Lt.T
    R -= np.tril(A) + np.tril(A, -1).T
    return float(np.linalg.norm(R) / np.linalg.norm(np.tril(A) + np.tril(A, -1).T))


# ---- tile layout helpers (W2:76-79, C2:280-309) --------------------------------------------------

def to_tiles(A: np.ndarray, b: int) -> dict:
    """Cut the lower triangle of A (N x N, N % b == 0) into b x b column-major tiles."""
    N = A.shape[0]
    assert N % b == 0
    nt = N // b
    return {(i, j): np.asfortranarray(A[i * b:(i + 1) * b, j * b:(j + 1) * b].copy()) for i in range(nt)
            for j in range(i + 1)}


def from_tiles(tiles: dict, nt: int, b: int) -> np.ndarray:
    """Assemble the lower tiles into an N x N array (strict upper tiles zero)."""
    A = np.zeros((nt * b, nt * b), dtype=np.float64, order="F")
    for (i, j), t in tiles.items():
        A[i * b:(i + 1) * b, j * b:(j + 1) * b] = t
    return A


# ---- the same path through OpenBLAS (scipy), i.e. the library the reference calls ---------------

def blas_potrf(A: np.ndarray, threads: int | None = None) -> tuple[np.ndarray, int]:
    """Monolithic LAPACK dpotrf('L') (V6:56 semantics; LP:54 is this call)."""
    from scipy.linalg import lapack
    with _threads(threads):
        c, info = lapack.dpotrf(A, lower=1, clean=0, overwrite_a=0)
    return c, int(info)


def blas_potrf_tiled(tiles: dict, nt: int, b: int, threads: int | None = 1) -> int:
    """Tile DAG (C1:278-333) with OpenBLAS tile kernels, in place on the dict of F-arrays.
    threads=1 mirrors the reference's OPENBLAS_NUM_THREADS=1 (benchmark.c:173-175)."""
    from scipy.linalg import blas, lapack
    with _threads(threads):
        for k in range(nt):
            c, info = lapack.dpotrf(tiles[(k, k)], lower=1, clean=0, overwrite_a=1)
            if c is not tiles[(k, k)]:
                tiles[(k, k)][...] = c
            if info:
                return k * b + int(info)
            Lkk = tiles[(k, k)]
            for i in range(k + 1, nt):
                x = blas.dtrsm(1.0, Lkk, tiles[(i, k)], side=1, lower=1, trans_a=1, diag=0, overwrite_b=1)
                if x is not tiles[(i, k)]:
                    tiles[(i, k)][...] = x
            for i in range(k + 1, nt):
                for j in range(k + 1, i + 1):
                    if i == j:
                        c = blas.dsyrk(-1.0, tiles[(i, k)], beta=1.0, c=tiles[(i, i)], trans=0, lower=1, overwrite_c=1)
                        if c is not tiles[(i, i)]:
                            tiles[(i, i)][...] = c
                    else:
                        c = blas.dgemm(-1.0, tiles[(i, k)], tiles[(j, k)], beta=1.0, c=tiles[(i, j)], trans_a=0,
                                       trans_b=1, overwrite_c=1)
                        if c is not tiles[(i, j)]:
                            tiles[(i, j)][...] = c
    return 0


class _threads:
    """Context manager: limit BLAS threads (threadpoolctl), None = leave alone."""

    def __init__(self, n):
        self.n, self.ctx = n, None

    def __enter__(self):
        if self.n is not None:
            from threadpoolctl import threadpool_limits
            self.ctx = threadpool_limits(limits=self.n, user_api="blas")
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def run_reference_binary(threads: int | None = None, dump_prefix: str | None = None,
                         skip_residual: bool = True) -> dict:
    """Run oracle/_ref/lapacke_dpotrf_ref (the reference's own CPU program, N=12000 fixed in its
    source, LP:22) and parse its stdout.  Returns {"N":…, "seconds":…, "gflops":…}."""
    if not os.path.exists(REF_BIN):
        raise FileNotFoundError(REF_BIN)
    env = dict(os.environ)
    if threads:
        env["OPENBLAS_NUM_THREADS"] = str(threads)
    if skip_residual:
        env["CHOL_REF_SKIP_RESIDUAL"] = "1"
    if dump_prefix:
        env["CHOL_REF_DUMP"] = dump_prefix
    out = subprocess.run([REF_BIN], env=env, capture_output=True, text=True, check=True).stdout
    res = {}
    for line in out.splitlines():
        if line.startswith("Taille N"):
            res["N"] = int(line.split("=")[1])
        elif line.startswith("Temps"):
            res["seconds"] = float(line.split("=")[1].split()[0])
        elif line.startswith("Performance"):
            res["gflops"] = float(line.split("=")[1].split()[0])
    return res
