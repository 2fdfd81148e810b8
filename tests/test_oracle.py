"""The CPU oracle against (a) the golden vectors produced by the reference's own CPU program
(oracle/_ref, pin_against_ref.py), (b) OpenBLAS — the library family the reference calls — and
(c) itself (tile DAG == monolithic).  No GPU."""
import hashlib
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_lapacke_dpotrf.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


def test_reference_input_reproduced(oracle, gold):
    A = oracle.lp_matrix(gold["N"])[: gold["m"], : gold["m"]]
    assert hashlib.sha256(np.asfortranarray(A).tobytes(order="F")).hexdigest() == gold["input_sha256_leading_block"]


@pytest.mark.parametrize("mode", ["monolithic", "tiled"])
def test_oracle_matches_reference_factor(oracle, gold, mode):
    """chol(A)[:m,:m] == chol(A[:m,:m]): replay the reference factor's leading block."""
    m = gold["m"]
    A = np.asfortranarray(oracle.lp_matrix(gold["N"])[:m, :m])
    if mode == "monolithic":
        L = A.copy(order="F")
        assert oracle.potrf_tile(L) == 0
    else:
        t = oracle.to_tiles(A, 256)
        assert oracle.potrf_tiled(t, m // 256, 256) == 0
        L = oracle.from_tiles(t, m // 256, 256)
    ref = np.array(gold["L"])
    got = L[np.array(gold["i"]), np.array(gold["j"])]
    scale = np.abs(np.array(gold["diag"])).max()
    assert np.abs(got - ref).max() <= 1e-13 * scale
    assert np.abs(np.diag(L) - np.array(gold["diag"])).max() <= 1e-13 * scale


@pytest.mark.parametrize("b", [1, 3, 16, 97, 256])
def test_tile_ops_match_openblas(oracle, b):
    from scipy.linalg import blas, lapack
    rng = np.random.default_rng(b)
    Ai = np.asfortranarray(rng.uniform(-0.5, 0.5, (b, b)))
    Aj = np.asfortranarray(rng.uniform(-0.5, 0.5, (b, b)))
    Cm = np.asfortranarray(rng.uniform(-0.5, 0.5, (b, b)))
    S = np.asfortranarray(Ai @ Ai.T + b * np.eye(b))
    tol = 1e-13
    # POTRF: lower factor, strict upper untouched
    L = S.copy(order="F")
    L[np.triu_indices(b, 1)] = 7.0
    assert oracle.potrf_tile(L) == 0
    Lr, info = lapack.dpotrf(S, lower=1, clean=1)
    assert info == 0 and np.abs(np.tril(L) - Lr).max() <= tol * np.abs(Lr).max()
    assert np.all(L[np.triu_indices(b, 1)] == 7.0)
    # TRSM
    X = Aj.copy(order="F")
    oracle.trsm_tile(np.asfortranarray(Lr), X)
    Xr = blas.dtrsm(1.0, Lr, Aj, side=1, lower=1, trans_a=1, diag=0)
    assert np.abs(X - Xr).max() <= tol * max(np.abs(Xr).max(), 1)
    # SYRK: lower only
    C1 = Cm.copy(order="F")
    oracle.syrk_tile(Ai, C1)
    full = Cm - Ai @ Ai.T
    assert np.abs(np.tril(C1) - np.tril(full)).max() <= tol * b
    assert np.all(C1[np.triu_indices(b, 1)] == Cm[np.triu_indices(b, 1)])
    # GEMM
    C2 = Cm.copy(order="F")
    oracle.gemm_tile(Ai, Aj, C2)
    assert np.abs(C2 - (Cm - Ai @ Aj.T)).max() <= tol * b


def test_potrf_info(oracle):
    S = np.asfortranarray(np.eye(8) * 4.0)
    S[5, 5] = -1.0
    assert oracle.potrf_tile(S) == 6
    S = np.asfortranarray(np.eye(8))
    S[2, 2] = np.nan
    assert oracle.potrf_tile(S) == 3


@pytest.mark.parametrize("N,b", [(256, 64), (384, 128), (96, 32)])
def test_tile_dag_equals_monolithic(oracle, N, b):
    A = oracle.plgsy(float(N), N, 42)
    L1 = A.copy(order="F")
    assert oracle.potrf_tile(L1) == 0
    t = oracle.to_tiles(A, b)
    assert oracle.potrf_tiled(t, N // b, b) == 0
    L2 = oracle.from_tiles(t, N // b, b)
    assert np.abs(np.tril(L1) - np.tril(L2)).max() <= 1e-13 * np.abs(L1).max()
    assert oracle.backward_error(A, L2) <= 1e-15 * 10
    # and through OpenBLAS tile kernels (the reference's library family)
    t2 = oracle.to_tiles(A, b)
    assert oracle.blas_potrf_tiled(t2, N // b, b) == 0
    assert np.abs(np.tril(oracle.from_tiles(t2, N // b, b)) - np.tril(L1)).max() <= 1e-13 * np.abs(L1).max()


def test_tile_dag_info_is_global_index(oracle):
    N, b = 128, 32
    A = oracle.plgsy(float(N), N, 1)
    A[70, 70] = -5.0
    t = oracle.to_tiles(A, b)
    assert oracle.potrf_tiled(t, N // b, b) == 71


def test_generator_c_equals_numpy(oracle):
    N = 77
    A = oracle.plgsy(float(N), N, 42)
    assert np.array_equal(A, oracle.plgsy_numpy(float(N), N, 42))
    assert np.array_equal(A, A.T)
    # tile view of the same matrix, with identity padding past the edge
    T = oracle.plgsy(float(N), N, 42, row0=64, col0=64, mb=32, nb=32)
    assert np.array_equal(T[:13, :13], A[64:, 64:])
    assert np.array_equal(T[13:, 13:], np.eye(19)) and not T[13:, :13].any() and not T[:13, 13:].any()
    # strictly diagonally dominant => SPD
    assert np.all(2 * np.abs(np.diag(A)) > np.abs(A).sum(1))


def test_backward_error_metric(oracle):
    from scipy.linalg import lapack
    A = oracle.plgsy(200.0, 200, 3)
    L, _ = lapack.dpotrf(A, lower=1, clean=1)
    L = np.asfortranarray(L)
    e = oracle.backward_error(A, L)
    assert 0 < e < 1e-15
    assert abs(e - oracle.backward_error_blas(A, L)) < 1e-16
    L[10, 3] += 1.0
    assert oracle.backward_error(A, L) > 1e-6
