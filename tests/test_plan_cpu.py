"""Invariants of the factorization plan (per-step task lists and panel pointer lists) on CPU tensors,
for several grids: the plan is the DAG the GPU executes, so it is checked against dag.build_dag."""
import numpy as np
import pytest
import torch

from dense_linear_app_b200 import dag
from dense_linear_app_b200.cholesky import TiledCholesky
from dense_linear_app_b200.grid import ProcessGrid
from dense_linear_app_b200.tiles import TileDesc, TileMatrix


class PlanOnly(TiledCholesky):
    """Builds the plan without touching CUDA or torch.distributed."""

    def __init__(self, A):
        self.A, self.nt, self.b = A, A.nt, A.b
        self.grid, self.rank, self.lay = A.grid, A.rank, A.layout
        self.dev, self.cuda, self.world = A.device, False, A.grid.size
        self.group, self.lookahead, self.nslots, self.transport, self.tr = None, True, 2, "nccl", None
        self.tile_bytes = self.b * self.b * 8
        self.panel = torch.empty((2, max(self.nt - 1, 1), self.b, self.b), dtype=torch.float64) if self.world > 1 else None
        self._build_plan()


@pytest.mark.parametrize("P,Q,N,b", [(1, 1, 80, 16), (1, 2, 96, 16), (2, 2, 112, 16), (2, 4, 176, 16), (2, 4, 100, 16)])
def test_plan_is_the_reference_dag(P, Q, N, b):
    g = ProcessGrid(P, Q)
    desc = TileDesc(b, b, b * b, N, N, 0, 0, N, N, P, Q)
    nt = (N + b - 1) // b
    want = {}
    for t in dag.build_dag(N, b):
        if t.op in ("SYRK", "GEMM"):
            want.setdefault(t.k, set()).add(t.out)
    seen = {k: set() for k in range(nt)}
    trsm_seen = {k: set() for k in range(nt)}
    for rank in range(g.size):
        M = TileMatrix(desc, rank, "cpu")
        pl = PlanOnly(M)
        base, tb = M.buf.data_ptr(), M.tile_bytes
        ptr2tile = {base + M.layout.index(i, j) * tb: (i, j) for i, j in M.layout.tiles()}
        lo, hi = base, base + M.buf.numel() * 8
        plo = pl.panel.data_ptr() if pl.panel is not None else 0
        phi = plo + pl.panel.numel() * 8 if pl.panel is not None else 0
        for k in range(nt):
            off, nd, na, ntot = pl.step_tasks[k]
            rec = pl.tasks_host[off:off + ntot]
            assert 0 <= nd <= 1 and nd <= na <= ntot
            for n_, (c, a, bb, flag) in enumerate(rec.tolist()):
                i, j = ptr2tile[c]                       # C is always an owned tile
                assert i >= j > k and (flag == 1) == (i == j)
                assert (i, j) not in seen[k]
                seen[k].add((i, j))
                assert (n_ < na) == (j == k + 1)          # part a = column k+1, first
                if n_ < nd:
                    assert (i, j) == (k + 1, k + 1)       # diagonal tile leads part a
                for p_, row in ((a, i), (bb, j)):         # operands: tile (row, k), local or in the receive slot
                    if lo <= p_ < hi:
                        assert ptr2tile[p_] == (row, k)
                    else:
                        assert plo <= p_ < phi and (p_ - plo) % tb == 0
                        assert (p_ - plo) // (pl.panel.stride(0) * 8) == k % 2
            if k == 0:
                # step 0, part b: one task per local tile of columns >= 2, in storage order; the upload
                # groups of factor_from_host tile that range without gaps
                head, groups = pl.step0_head, pl.step0_groups
                assert [ptr2tile[c] for c in rec[na:, 0].tolist()] == [t for t in M.layout.tiles() if t[1] >= 2]
                assert rec[na:, 0].tolist() == [base + (head + n_) * tb for n_ in range(ntot - na)]
                t_next, lo_next = 0, head
                for t0, t1, tlo, thi in groups:
                    assert (t0, tlo) == (t_next, lo_next) and t1 - t0 == thi - tlo > 0
                    t_next, lo_next = t1, thi
                assert t_next == ntot - na and lo_next == M.layout.ntiles
            toff, cnt = pl.step_trsm[k]
            for p_ in pl.d_trsm_ptrs[toff:toff + cnt].tolist():
                i, j = ptr2tile[p_]
                assert j == k and i > k
                trsm_seen[k].add(i)
    for k in range(nt):
        assert seen[k] == want.get(k, set()), k           # every update of wave k exactly once over all ranks
        assert trsm_seen[k] == set(range(k + 1, nt))


def test_tilematrix_roundtrip_and_identity_padding():
    N, b = 50, 16
    rng = np.random.default_rng(0)
    A = rng.standard_normal((N, N))
    A = A + A.T
    M = TileMatrix(TileDesc.square(N, b), 0, "cpu").from_numpy(A)
    back = M.to_numpy()
    assert np.array_equal(np.tril(back), np.tril(A))
    last = M.tile(3, 3).numpy().T                         # column-major view of the ragged diagonal tile
    assert np.array_equal(last[2:, 2:], np.eye(14)) and not last[2:, :2].any()
    assert M.clone().buf.data_ptr() != M.buf.data_ptr()


@pytest.mark.parametrize("N,b", [(16 * 9, 16), (16 * 24, 16), (16 * 64, 16), (16 * 3, 16)])
def test_lazy_plan_of_the_host_resident_path(N, b):
    """factor_from_host on one rank applies the first S steps lazily per upload group (cholesky._build_lazy_plan):
    the part-b tasks of steps 1..S-1 are cut by upload group without loss or duplication, and everything the panel
    chain of those steps needs (columns <= S) is in the head or in group 0."""
    M = TileMatrix(TileDesc.square(N, b), 0, "cpu")
    pl = PlanOnly(M)
    nt, S = M.nt, pl.lazy_steps
    base, tb = M.buf.data_ptr(), M.tile_bytes
    if nt < 6:
        assert S == 1 and pl.d_lazy is None            # too small: nothing to do lazily
        return
    assert 2 <= S <= 4
    lazy = pl.d_lazy.numpy()
    bounds = [(g[2], g[3]) for g in pl.step0_groups]
    ptr2tile = {base + M.layout.index(i, j) * tb: (i, j) for i, j in M.layout.tiles()}
    early_hi = bounds[0][1]
    for j in range(S + 1):                              # every column <= S is on the device before group 1 arrives
        assert M.layout.index(nt - 1, j) < early_hi
    for k in range(1, S):
        off, nd, na, ntot = pl.step_tasks[k]
        want = {tuple(r) for r in pl.tasks_host[off + na:off + ntot].tolist()}
        got = []
        for g, (lo, hi) in enumerate(bounds):
            o, cnt = pl.lazy_index[(k, g)]
            for c, a, bb, flag in lazy[o:o + cnt].tolist():
                t = (c - base) // tb
                assert lo <= t < hi                     # the task's C tile belongs to this upload group
                i, j = ptr2tile[c]
                assert j >= k + 2                       # part b only: column k+1 is part a
                got.append((c, a, bb, flag))
        assert len(got) == len(want) and set(got) == want
