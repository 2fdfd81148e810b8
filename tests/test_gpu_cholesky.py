"""Whole-matrix path on the GPU: TiledCholesky / potrf_tile_desc (replaces CHAMELEON_dpotrf_Tile,
v6_test.c:56) against the oracle's tile DAG and monolithic dpotrf; generator parity; batched
POTRF; worker Execute; the v6_test command line.  Gates (north_star): backward error
||A - L L^T||_F/||A||_F <= 1e-13; |L - Lref| <= 1e-10 * max(|Lref|, 1e-3 * max|Lref|)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ill_conditioned_spd(n, cond, kind, seed=0):
    """SPD test matrices far from cond ~ 1 (the generators of the reference only make those):
    geo      Q diag(logspace(0, -log10 cond)) Q^T        geometrically spread spectrum
    cluster  Q diag(1, ..., 1, 1/cond x n/8) Q^T           one small cluster
    graded   D B D, B well conditioned, D = diag(logspace)   badly scaled rows/columns"""
    rng = np.random.default_rng(seed)
    if kind == "graded":
        Bm = rng.standard_normal((n, n))
        Bm = Bm @ Bm.T / n + np.eye(n)
        s = np.logspace(0, -np.log10(cond) / 2, n)
        A = (Bm * s).T * s
    else:
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        d = np.logspace(0, -np.log10(cond), n) if kind == "geo" else np.where(np.arange(n) >= n - n // 8, 1.0 / cond, 1.0)
        A = (Q * d) @ Q.T
    return np.asfortranarray((A + A.T) / 2)


def element_gate(L, Lref):
    floor = 1e-3 * np.abs(Lref).max()
    return np.all(np.abs(L - Lref) <= 1e-10 * np.maximum(np.abs(Lref), floor))


@pytest.mark.parametrize("N,b", [(64, 16), (96, 32), (256, 64), (1000, 128), (1024, 256), (2048, 512)])
def test_generator_matches_oracle_bit_exact(cuda_lib, oracle, N, b):
    from dense_linear_app_b200.tiles import TileDesc, TileMatrix
    M = TileMatrix(TileDesc.square(N, b)).generate(float(N), 42)
    got = M.to_numpy()
    ref = oracle.plgsy(float(N), N, 42) if N <= 1024 else oracle.plgsy_numpy(float(N), N, 42)
    assert np.array_equal(np.tril(got), np.tril(ref))
    # diagonal tiles are generated whole (mirrored upper part)
    assert np.array_equal(got[:b, :b], ref[:b, :b])


@pytest.mark.parametrize("N,b,lookahead", [(4, 4, True), (12, 4, True), (64, 16, False), (200, 64, True),
                                           (1000, 128, True), (1024, 256, False), (2048, 512, True),
                                           (4096, 512, True)])
def test_factor_matches_oracle(cuda_lib, oracle, N, b, lookahead):
    """configs[0] of BASELINE.json is the last case: N=4096, tile 512."""
    from scipy.linalg import lapack
    from dense_linear_app_b200.cholesky import TiledCholesky
    from dense_linear_app_b200.tiles import TileDesc, TileMatrix
    A = oracle.plgsy(float(N), N, 42) if N <= 1024 else oracle.plgsy_numpy(float(N), N, 42)
    M = TileMatrix(TileDesc.square(N, b)).from_numpy(A)
    M0 = M.clone()
    ch = TiledCholesky(M, lookahead=lookahead)
    ch.factor()
    assert ch.info() == 0
    L = np.tril(M.to_numpy())
    if N <= 256 and N % b == 0:
        t = oracle.to_tiles(A, b)                      # plain-C restatement of the same tile DAG
        assert oracle.potrf_tiled(t, N // b, b) == 0
        Lref = np.tril(oracle.from_tiles(t, N // b, b))
    else:
        Lref, info = lapack.dpotrf(A, lower=1, clean=1)   # OpenBLAS: the reference's library
        assert info == 0
    assert element_gate(L, Lref)
    assert np.abs(L - Lref).max() <= 1e-13 * np.abs(Lref).max()
    bwd = np.linalg.norm(L @ L.T - A) / np.linalg.norm(A)
    assert bwd <= 1e-13
    res = ch.residual(M0)                               # device-side residual agrees with the host one
    assert res["fro"] <= 1e-13 and abs(res["fro"] - bwd) <= 2e-16 + 0.2 * bwd
    assert res["inf"] <= 1e-13


def test_potrf_tile_desc_entry_point_and_info(cuda_lib, oracle):
    from dense_linear_app_b200.cholesky import potrf_tile_desc
    from dense_linear_app_b200.tiles import TileDesc, TileMatrix
    N, b = 512, 128
    A = oracle.plgsy(float(N), N, 1)
    assert potrf_tile_desc("L", TileMatrix(TileDesc.square(N, b)).from_numpy(A)) == 0
    A[300, 300] = -4.0
    assert potrf_tile_desc("L", TileMatrix(TileDesc.square(N, b)).from_numpy(A)) == 301   # global LAPACK index
    # uplo = 'U' on the same (symmetric) input: the bad pivot is found at the same index
    assert potrf_tile_desc("U", TileMatrix(TileDesc.square(N, b)).from_numpy(A)) == 301
    with pytest.raises(ValueError):
        potrf_tile_desc("B", TileMatrix(TileDesc.square(N, b)).from_numpy(A))   # ChamUpperLower: not a Cholesky


def test_strict_upper_of_diagonal_tiles_untouched(cuda_lib, oracle):
    from dense_linear_app_b200.cholesky import TiledCholesky
    from dense_linear_app_b200.tiles import TileDesc, TileMatrix
    N, b = 512, 128
    M = TileMatrix(TileDesc.square(N, b)).generate(float(N), 9)
    before = M.to_numpy()
    ch = TiledCholesky(M)
    ch.factor()
    assert ch.info() == 0
    after = M.to_numpy()
    for k in range(N // b):
        s = slice(k * b, (k + 1) * b)
        assert np.array_equal(np.triu(after[s, s], 1), np.triu(before[s, s], 1))


def test_config2_properties_full_size(cuda_lib):
    """BASELINE configs[1] (N=16384, tile 1024): too big for a CPU factor inside the test budget, so
    check size-independent properties: info == 0, device backward error <= 1e-13, and the leading
    2048 block equals the oracle-checked factor of the leading block (chol(A)[:m,:m] = chol(A[:m,:m]))."""
    from scipy.linalg import lapack
    from dense_linear_app_b200.cholesky import TiledCholesky
    from dense_linear_app_b200.tiles import TileDesc, TileMatrix
    N, b, m = 16384, 1024, 2048
    M = TileMatrix(TileDesc.square(N, b)).generate(float(N), 42)
    lead = np.zeros((m, m))
    for i in range(m // b):
        for j in range(i + 1):
            lead[i * b:(i + 1) * b, j * b:(j + 1) * b] = M.tile(i, j).cpu().numpy().T
    lead = np.tril(lead) + np.tril(lead, -1).T
    M0 = M.clone()
    ch = TiledCholesky(M)
    ch.factor()
    assert ch.info() == 0
    Lref, info = lapack.dpotrf(lead, lower=1, clean=1)
    got = np.zeros((m, m))
    for i in range(m // b):
        for j in range(i + 1):
            got[i * b:(i + 1) * b, j * b:(j + 1) * b] = M.tile(i, j).cpu().numpy().T
    assert element_gate(np.tril(got), Lref)
    res = ch.residual(M0)
    assert res["fro"] <= 1e-13 and res["inf"] <= 1e-13


@pytest.mark.parametrize("kind", ["geo", "cluster", "graded"])
@pytest.mark.parametrize("cond", [1e4, 1e8, 1e10])
def test_factor_ill_conditioned_backward_error(cuda_lib, kind, cond):
    """north_star limits the ELEMENT gate to well-conditioned inputs, not the backward-error gate:
    ||A - L L^T||_F / ||A||_F <= 1e-13 must survive cond 1e4 ... 1e10 although TRSM and the POTRF panel
    multiply by explicitly inverted 128 x 128 diagonal blocks."""
    from scipy.linalg import lapack
    from dense_linear_app_b200.cholesky import TiledCholesky
    from dense_linear_app_b200.tiles import TileDesc, TileMatrix
    N, b = 2048, 512
    A = ill_conditioned_spd(N, cond, kind)
    M = TileMatrix(TileDesc.square(N, b)).from_numpy(A)
    M0 = M.clone()
    ch = TiledCholesky(M)
    ch.factor()
    assert ch.info() == 0
    L = np.tril(M.to_numpy())
    bwd = np.linalg.norm(L @ L.T - A) / np.linalg.norm(A)
    Lref, info = lapack.dpotrf(A, lower=1, clean=1)
    assert info == 0
    bwd_ref = np.linalg.norm(Lref @ Lref.T - A) / np.linalg.norm(A)
    assert bwd <= 1e-13, (bwd, bwd_ref)
    assert bwd <= 50 * max(bwd_ref, 1e-16)             # and not qualitatively worse than LAPACK itself
    assert ch.residual(M0)["fro"] <= 1e-13
    # forward error scales with the condition number (as LAPACK's own does): only a sanity bound here
    assert np.abs(L - Lref).max() <= 1e-14 * cond * np.abs(Lref).max()


@pytest.mark.parametrize("b", [256, 1024])
@pytest.mark.parametrize("cond", [1e6, 1e10])
def test_tile_ops_ill_conditioned(cuda_lib, b, cond):
    """POTRF and TRSM tile ops on an ill-conditioned tile: residuals ||A - L L^T|| and ||X L^T - B||
    relative to the data stay at rounding level."""
    from dense_linear_app_b200 import tile_ops
    A = ill_conditioned_spd(b, cond, "geo", seed=3)
    dA = torch.from_numpy(np.ascontiguousarray(A.T)).cuda()
    info = tile_ops.potrf_tile(dA)
    assert int(info.item()) == 0
    L = np.tril(dA.cpu().numpy().T)
    assert np.linalg.norm(L @ L.T - A) / np.linalg.norm(A) <= 1e-13
    rng = np.random.default_rng(5)
    Bm = rng.standard_normal((b, b))
    dB = torch.from_numpy(np.ascontiguousarray(Bm.T)).cuda()
    tile_ops.trsm_tile(dA, dB)
    X = dB.cpu().numpy().T
    # normwise backward error of the solve X L^T = B
    assert np.linalg.norm(X @ L.T - Bm) / (np.linalg.norm(X) * np.linalg.norm(L) + np.linalg.norm(Bm)) <= 1e-13


def test_gpu_factor_matches_the_reference_programs_own_factor(cuda_lib, oracle):
    """The one fixture that comes from the REFERENCE itself (tests/golden/ref_lapacke_dpotrf.json: output of
    its lapacke_dpotrf.c, N=12000, compiled from /root/reference by oracle/Makefile): the GPU factors the
    leading block of the same input (chol(A)[:m,:m] == chol(A[:m,:m])) and is compared with the recorded
    entries directly, not through the oracle."""
    import hashlib
    from dense_linear_app_b200.cholesky import TiledCholesky
    from dense_linear_app_b200.tiles import TileDesc, TileMatrix
    with open(os.path.join(ROOT, "tests", "golden", "ref_lapacke_dpotrf.json")) as f:
        gold = json.load(f)
    m = gold["m"]
    A = np.asfortranarray(oracle.lp_matrix(gold["N"])[:m, :m])       # input generator only (sha-checked below)
    assert hashlib.sha256(A.tobytes(order="F")).hexdigest() == gold["input_sha256_leading_block"]
    for b in (256, 512):
        M = TileMatrix(TileDesc.square(m, b)).from_numpy(A)
        ch = TiledCholesky(M)
        ch.factor()
        assert ch.info() == 0
        L = np.tril(M.to_numpy())
        scale = np.abs(np.array(gold["diag"])).max()
        got = L[np.array(gold["i"]), np.array(gold["j"])]
        assert np.abs(got - np.array(gold["L"])).max() <= 1e-13 * scale
        assert np.abs(np.diag(L) - np.array(gold["diag"])).max() <= 1e-13 * scale
        ref = np.array(gold["L"])
        floor = 1e-3 * scale
        assert np.all(np.abs(got - ref) <= 1e-10 * np.maximum(np.abs(ref), floor))


@pytest.mark.parametrize("n,batch", [(1, 3), (5, 7), (32, 64), (64, 20), (96, 11), (128, 33), (160, 9), (200, 5),
                                     (224, 6), (256, 40), (256, 700), (288, 4)])
def test_potrf_batched(cuda_lib, oracle, n, batch):
    """configs[4] shape (many small SPD matrices, n=256) at a test-sized batch."""
    from dense_linear_app_b200 import tile_ops
    mats = [oracle.plgsy(float(n), n, 42 + i) for i in range(batch)]
    mats[batch // 2] = mats[batch // 2].copy()
    bad = n // 2
    mats[batch // 2][bad, bad] = -1.0
    d = torch.from_numpy(np.stack([np.ascontiguousarray(m.T) for m in mats])).cuda()
    info = tile_ops.potrf_batched(d).cpu().numpy()
    out = d.cpu().numpy()
    for i, m in enumerate(mats):
        ref = m.copy(order="F")
        want = oracle.potrf_tile(ref)
        assert info[i] == want
        if want == 0:
            got = out[i].T
            assert np.abs(np.tril(got) - np.tril(ref)).max() <= 1e-13 * np.abs(ref).max()
            assert np.array_equal(np.triu(got, 1), np.triu(m, 1))


@pytest.mark.parametrize("n,lda,pad", [(256, 258, 6), (96, 96, 0), (64, 65, 1), (128, 130, 2)])
def test_potrf_batched_strided_through_the_c_abi(cuda_lib, oracle, n, lda, pad):
    """lda > n and a padded matrix stride: even lda/stride take the left-looking DMMA kernel, odd ones
    the generic kernels; rows n..lda-1 and the padding must come back untouched."""
    batch = 7
    stride = lda * n + pad
    host = np.full(batch * stride, -7.5)
    mats = []
    for i in range(batch):
        m = oracle.plgsy(float(n), n, 100 + i)
        mats.append(m)
        blk = host[i * stride:i * stride + lda * n].reshape(n, lda)      # [col, row]
        blk[:, :n] = m.T
    d = torch.from_numpy(host.copy()).cuda()
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    cuda_lib.call("chol_potrf_batched", n, batch, d.data_ptr(), lda, stride, info.data_ptr(),
                  torch.cuda.current_stream().cuda_stream)
    out = d.cpu().numpy()
    assert not info.cpu().numpy().any()
    for i, m in enumerate(mats):
        ref = m.copy(order="F")
        assert oracle.potrf_tile(ref) == 0
        blk = out[i * stride:i * stride + lda * n].reshape(n, lda)
        got = blk[:, :n].T
        assert np.abs(np.tril(got) - np.tril(ref)).max() <= 1e-13 * np.abs(ref).max()
        assert np.array_equal(np.triu(got, 1), np.triu(m, 1))
        assert np.all(blk[:, n:] == -7.5)
        assert np.all(out[i * stride + lda * n:(i + 1) * stride] == -7.5)


@pytest.mark.parametrize("N,b", [(1024, 256), (1000, 128)])
def test_potrf_tile_desc_upper_is_the_transposed_lower_factor(cuda_lib, N, b):
    """uplo = 'U' (ChamUpper of CHAMELEON_dpotrf_Tile, v6_test.c:56; --uplo U of the v3 driver): storage
    position (i, j) holds the upper tile (j, i); the result must be U(j, i) = L(i, j)^T bit for bit, and the
    strict LOWER triangle of the diagonal tiles must not be touched."""
    from dense_linear_app_b200.cholesky import potrf_tile_desc, transpose_tiles
    from dense_linear_app_b200.tiles import TileDesc, TileMatrix
    AL = TileMatrix(TileDesc.square(N, b)).generate(float(N), 3)
    AU = AL.clone()
    transpose_tiles(AU)                                  # now the upper tiles
    nt = AL.nt
    for k in range(nt):                                  # sentinel in the strict lower triangle of the diagonal tiles
        t = AU.tile(k, k)                                # torch view [col][row]: strict lower = row > col = triu(t, 1)
        t += torch.triu(torch.full_like(t, 123.0), 1)
    assert potrf_tile_desc("L", AL) == 0
    assert potrf_tile_desc("U", AU) == 0
    torch.cuda.synchronize()
    for i, j in AL.layout.tiles():
        l, u = AL.tile(i, j), AU.tile(i, j)
        if i == j:
            assert torch.equal(torch.tril(u), torch.triu(l).T.contiguous())      # upper triangle of U = (lower of L)^T
            strict = torch.triu(torch.ones_like(u, dtype=torch.bool), 1)
            assert bool((u[strict] >= 100.0).all())                               # sentinel region untouched
        else:
            assert torch.equal(u, l.T.contiguous())


def test_potrf_batched_from_host_matches_device_path(cuda_lib, oracle):
    """Host-resident batch through the three-stream pipeline == the device path, info included."""
    from dense_linear_app_b200 import tile_ops
    n, batch = 96, 53
    mats = [oracle.plgsy(float(n), n, 7 + i) for i in range(batch)]
    mats[11] = mats[11].copy()
    mats[11][40, 40] = -3.0
    host = torch.from_numpy(np.stack([np.ascontiguousarray(m.T) for m in mats])).pin_memory()
    out = torch.empty_like(host).pin_memory()
    hinfo = torch.empty(batch, dtype=torch.int32).pin_memory()
    for chunks in (1, 4, 16, 100):
        out.zero_()
        tile_ops.potrf_batched_from_host(host, out, hinfo, chunks=chunks)
        torch.cuda.current_stream().synchronize()
        d = host.cuda()
        info = tile_ops.potrf_batched(d)
        # bit patterns, not values: the matrix with the bad pivot holds NaNs past the failure
        assert torch.equal(out.cuda().view(torch.int64), d.view(torch.int64)) and torch.equal(hinfo.cuda(), info)
        assert int(hinfo[11]) == 41 and int((hinfo != 0).sum()) == 1


def test_potrf_batched_config4_full_size_properties(cuda_lib):
    """BASELINE configs[4] at its full size (10 000 x 256): info == 0 everywhere and the backward
    error ||A - L L^T||_F / ||A||_F of EVERY matrix <= 1e-13 (checked with torch.bmm, the checker)."""
    from dense_linear_app_b200 import tile_ops
    batch, n = 10000, 256
    st = torch.cuda.current_stream().cuda_stream
    A = torch.empty(batch, n, n, dtype=torch.float64, device="cuda")
    for i in range(batch):
        cuda_lib.call("chol_plgsy_tile", float(n), n, n, A[i].data_ptr(), n, n, 0, 0, n, 42 + i, st)
    A0 = A.clone()
    info = tile_ops.potrf_batched(A)
    assert int((info != 0).sum().item()) == 0
    Lt = torch.triu(A)                      # torch sees the transpose: upper == L^T
    full = torch.triu(A0) + torch.triu(A0, 1).transpose(1, 2)
    R = full - Lt.transpose(1, 2) @ Lt
    err = torch.linalg.matrix_norm(R) / torch.linalg.matrix_norm(full)
    assert float(err.max().item()) <= 1e-13
    assert torch.equal(torch.tril(A, -1), torch.tril(A0, -1))      # strict upper (col-major) untouched


def test_worker_execute_runs_the_client_dag(cuda_lib, oracle):
    """ArmoniK path end to end: client wave loop (dag.run_waves) -> JSON payload + host blobs ->
    DagCholeskyWorker.Execute on the GPU -> blobs; result == dpotrf."""
    from dense_linear_app_b200 import dag, worker
    N, B = 48, 16
    A = dag.enforce_strict_diag_dominance(dag.make_spd_like_chameleon(N))
    nb = N // B
    blocks = {dag.block_id_from_ij(i, j): dag.extract_block(A, B, i, j).tobytes(order="F")
              for i in range(nb) for j in range(i + 1)}
    out = dag.run_waves(N, B, blocks, worker.execute)
    L = np.zeros((N, N))
    for i in range(nb):
        for j in range(i + 1):
            L[i * B:(i + 1) * B, j * B:(j + 1) * B] = np.frombuffer(out[dag.block_id_from_ij(i, j)]).reshape(B, B).T
    L = np.tril(L)
    ref = A.copy(order="F")
    assert oracle.potrf_tile(ref) == 0
    assert np.abs(L - np.tril(ref)).max() <= 1e-13 * np.abs(ref).max()


def test_worker_runs_the_v1_client_input_with_nonsymmetric_diagonal_tiles(cuda_lib, oracle):
    """v1 client input (C1:102-108,189-192): independent N(0, 0.1^2) tiles, +B on the diagonal tiles'
    diagonal — the diagonal tiles are NOT symmetric, so this passes only if the GPU POTRF and SYRK read
    and write the lower triangle alone.  Result == LAPACK factor of the implied symmetric matrix, and the
    strict upper triangles of the diagonal tiles come back bit for bit."""
    from dense_linear_app_b200 import dag, worker
    N, B = 192, 64
    blocks = dag.make_blocks_v1(N, B)
    nb = N // B
    A = np.zeros((N, N))
    for (i, j), blob in blocks.items():
        t = blob.reshape(B, B).T
        A[i * B:(i + 1) * B, j * B:(j + 1) * B] = np.tril(t) if i == j else t
    A = A + np.tril(A, -1).T
    named = {dag.block_id_from_ij(i, j): v.tobytes() for (i, j), v in blocks.items()}
    out = dag.run_waves(N, B, named, worker.execute)
    L = np.zeros((N, N))
    for i in range(nb):
        for j in range(i + 1):
            t = np.frombuffer(out[dag.block_id_from_ij(i, j)]).reshape(B, B).T
            L[i * B:(i + 1) * B, j * B:(j + 1) * B] = t
            if i == j:
                assert np.array_equal(np.triu(t, 1), np.triu(blocks[(i, j)].reshape(B, B).T, 1))
    L = np.tril(L)
    ref = A.copy(order="F")
    assert oracle.potrf_tile(ref) == 0
    assert np.abs(L - np.tril(ref)).max() <= 1e-13 * np.abs(ref).max()
    assert oracle.backward_error(np.asfortranarray(A), np.asfortranarray(L)) <= 1e-13


def test_worker_error_statuses(cuda_lib):
    """Same status texts as worker_distrib.cpp:194-195,218-220,243-244,547-549,558-560."""
    from dense_linear_app_b200 import worker
    w = worker.DagCholeskyWorker()
    B = 4
    good = np.eye(B).tobytes()
    st = w.Execute(worker.TaskHandler(json.dumps({"op": "POTRF", "B": B, "in": "x"}), {}))
    assert not st.ok and st.details == "[Worker][POTF] Missing dependency: x"
    st = w.Execute(worker.TaskHandler(json.dumps({"op": "GEMM", "B": B, "inC": "c", "inAi": "a", "inAj": "b"}),
                                      {"c": good, "a": good}))
    assert st.details == "[Worker][GEMM] Missing dependency: b"
    st = w.Execute(worker.TaskHandler(json.dumps({"op": "POTRF", "B": B, "in": "x"}), {"x": good[:-8]}))
    assert st.details == "[Worker][POTF] Bad block size: expected 16 doubles, got 15"
    st = w.Execute(worker.TaskHandler(json.dumps({"op": "LU", "B": B}), {}))
    assert st.details == "Unknown op=LU"
    notpd = -np.eye(B)
    st = w.Execute(worker.TaskHandler(json.dumps({"op": "POTRF", "B": B, "in": "x"}), {"x": notpd.tobytes()}))
    assert st.details == "Exception: [Worker][POTF] dpotrf info=1"
    st = w.Execute(worker.TaskHandler("{not json", {}))
    assert not st.ok and st.details.startswith("Exception: ")
    th = worker.TaskHandler(json.dumps({"op": "POTRF", "B": B, "in": "x"}), {"x": (4 * np.eye(B)).tobytes()},
                            expected_results=["out-1"])
    assert w.Execute(th).ok and np.array_equal(np.frombuffer(th.results["out-1"]).reshape(B, B), 2 * np.eye(B))


def test_v6_test_command_line(cuda_lib):
    """The driver contract benchmark.c relies on: 16 positionals, the two parsed stdout lines, exit 0."""
    from dense_linear_app_b200 import bench_sweep
    args = bench_sweep.driver_argv(0, 1, 1000, 128, 1, 1, 42)
    pr = subprocess.run([sys.executable, "-m", "dense_linear_app_b200.v6_test"] + args, cwd=ROOT,
                        capture_output=True, text=True, timeout=600)
    assert pr.returncode == 0, pr.stderr
    gflops, rel = bench_sweep.parse_metrics(pr.stdout)
    assert gflops > 0 and 0 <= rel < 1e-10
    assert "[setup] ncpu=0 ngpu=1 N=1000 NB=128" in pr.stdout and "N = 1000, NB = 128" in pr.stdout
    assert "PASS" in pr.stdout


@pytest.mark.parametrize("N,b", [(256, 128), (1024, 128), (2048, 256), (1000, 128), (4096, 128), (3072, 64)])
def test_factor_from_host_matches_device_path(cuda_lib, oracle, N, b):
    """End-to-end entry (pinned host tiles -> H2D pipelined under the first steps, which are applied lazily to
    the column groups as they arrive -> factor -> D2H of each finished column): bit-identical to the
    device-resident path, input buffer left untouched."""
    from dense_linear_app_b200.cholesky import TiledCholesky
    from dense_linear_app_b200.tiles import TileDesc, TileMatrix
    M = TileMatrix(TileDesc.square(N, b)).generate(float(N), 3)
    pristine = M.buf.clone()
    ch = TiledCholesky(M)
    ch.factor()
    assert ch.info() == 0
    want = M.buf.cpu()
    hin = torch.empty(M.buf.shape, dtype=torch.float64).pin_memory()
    hin.copy_(pristine)
    hout = torch.zeros(M.buf.shape, dtype=torch.float64).pin_memory()
    M.buf.zero_()                                   # the device buffer must be filled by the upload
    if N // b >= 16:
        assert ch.lazy_steps >= 2                   # the lazy first steps are what is being tested
    for _ in range(2):                              # twice: stream/event reuse across calls
        ch.factor_from_host(hin, hout)
        torch.cuda.synchronize()
        assert ch.info() == 0
        assert torch.equal(hout, want)
        assert torch.equal(hin, pristine.cpu())
        hout.zero_()
