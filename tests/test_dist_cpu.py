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


def _worker(rank, world, port, P, Q, N, b, lookahead, bad, transport="nccl"):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CHOL_PANEL_TRANSPORT=transport)
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
        assert ch.transport == transport and (ch.tr is not None) == (transport == "peer")
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
        if transport == "peer":
            # a second factorization through the same slots (next epoch of the flag counters), and
            # the volume: every tile reaches P+Q-2 readers, not all P*Q-1
            M.buf.copy_(TileMatrix(desc, rank, "cpu").from_numpy(A).buf)
            ch.tr.pushed_bytes = 0
            ch.factor()
            assert ch.info() == 0
            got2 = M.to_numpy()
            assert np.array_equal(got2, got)
            sent = torch.tensor([ch.tr.pushed_bytes], dtype=torch.int64)
            recv = torch.tensor([ch.tr.bytes_received_per_run()], dtype=torch.int64)
            dist.all_reduce(sent)
            dist.all_reduce(recv)
            assert int(sent) == int(recv)
            tb = b * b * 8
            ntl = nt * (nt - 1) // 2
            diag_msgs = sum(len(ch.tr.diag_readers(k)) for k in range(nt - 1)) if P > 1 else 0
            assert int(sent) <= ntl * (P + Q - 2) * tb + diag_msgs * (tb + ch.tr.work_bytes)
            ch.close()
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


@pytest.mark.parametrize("P,Q,N,b,lookahead", [
    (1, 2, 96, 16, True),
    (2, 2, 116, 16, True),      # ragged edge
    (2, 1, 80, 16, False),
    (2, 4, 208, 16, True),      # the 8-GPU grid shape
    (3, 2, 112, 16, True),      # P does not divide Q: the strided subsets are not "every other tile"
])
def test_peer_push_transport_shared_memory(P, Q, N, b, lookahead):
    """The copy-engine transport's host logic (slots, reader subsets, credits, flag epochs) with
    shared memory standing in for the IPC-mapped peer buffers."""
    world = P * Q
    mp.spawn(_worker, args=(world, _free_port(), P, Q, N, b, lookahead, None, "peer"), nprocs=world, join=True)


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
