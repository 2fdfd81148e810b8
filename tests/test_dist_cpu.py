"""The multi-rank schedule (2D block-cyclic ownership, L_kk / panel broadcasts, per-rank task
lists, residual) on CPU tensors over gloo, with the tile kernels played by the oracle
(tests/_cpu_backend.py).  Checks that every rank ends up with exactly its tiles of dpotrf(A)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, P, Q, N, b, lookahead, bad):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from _cpu_backend import OracleBackedCholesky
        from dense_linear_app_b200.tiles import TileDesc, TileMatrix
        from oracle import oracle as O
        A = O.plgsy(float(N), N, 42)
        if bad is not None:
            A[bad, bad] = -1.0
        desc = TileDesc(b, b, b * b, N, N, 0, 0, N, N, P, Q)
        M = TileMatrix(desc, rank, "cpu").from_numpy(A)
        M0 = M.clone()
        ch = OracleBackedCholesky(M, lookahead=lookahead)
        ch.factor()
        info = ch.info()
        if bad is not None:
            assert info == bad + 1, (info, bad)
            return
        assert info == 0
        from scipy.linalg import lapack
        nt = M.nt
        Apad = np.eye(nt * b)
        Apad[:N, :N] = A
        Lref, _ = lapack.dpotrf(Apad, lower=1, clean=1)
        got = M.to_numpy()
        for i, j in M.layout.tiles():
            r0, c0 = i * b, j * b
            blk = got[r0:min(N, r0 + b), c0:min(N, c0 + b)]
            ref = Lref[r0:min(N, r0 + b), c0:min(N, c0 + b)]
            if i == j:
                blk, ref = np.tril(blk), np.tril(ref)
            assert np.abs(blk - ref).max() <= 1e-13 * np.abs(Lref).max(), (rank, i, j)
        res = ch.residual(M0)
        assert res["fro"] < 1e-15 * 10 and res["inf"] < 1e-14, res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P,Q,N,b,lookahead", [
    (1, 2, 96, 16, True),
    (1, 2, 100, 16, False),     # ragged edge
    (2, 2, 112, 16, True),
    (2, 1, 80, 16, True),
    (2, 4, 160, 16, True),      # the 8-GPU grid shape
])
def test_block_cyclic_schedule_gloo(P, Q, N, b, lookahead):
    world = P * Q
    mp.spawn(_worker, args=(world, _free_port(), P, Q, N, b, lookahead, None), nprocs=world, join=True)


def test_info_reduced_over_ranks_gloo():
    mp.spawn(_worker, args=(2, _free_port(), 1, 2, 64, 16, True, 37), nprocs=2, join=True)


def test_single_rank_cpu_double(oracle):
    """World of one: same schedule code path the GPU uses, kernels played by the oracle."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _cpu_backend import OracleBackedCholesky
    from dense_linear_app_b200.tiles import TileDesc, TileMatrix
    N, b = 90, 16
    A = oracle.plgsy(float(N), N, 7)
    M = TileMatrix(TileDesc.square(N, b), 0, "cpu").from_numpy(A)
    M0 = M.clone()
    ch = OracleBackedCholesky(M)
    ch.factor()
    assert ch.info() == 0
    L = np.tril(M.to_numpy())
    assert oracle.backward_error(A, np.asfortranarray(L)) < 1e-15 * 10
    assert ch.residual(M0)["fro"] < 1e-15 * 10
    assert ch.n_update_tasks == sum(1 for t in __import__("dense_linear_app_b200.dag", fromlist=["x"]).build_dag(N, b)
                                    if t.op in ("SYRK", "GEMM"))
