"""Race check of the stream/event schedule, on the CPU.

The factorization enqueues kernels on two CUDA streams tied together by events (lookahead).  A
missing event wait would be an intermittent data race on the GPU.  Here the real ``TiledCholesky._run``
(and ``residual``) is executed with fake streams/events that carry vector clocks, the kernel entry
points record which tiles they read and write, and every conflicting pair of operations (write/write
or read/write on the same tile or scratch buffer) must be ordered by happens-before — for one rank
of several grids, with and without lookahead.  No CUDA, no communication: broadcasts are recorded as
a read (owner) or a write (receiver) of the tiles they carry.
"""
import contextlib
import ctypes as C

import numpy as np
import pytest
import torch

from dense_linear_app_b200 import cholesky as chol_mod
from dense_linear_app_b200.cholesky import TiledCholesky
from dense_linear_app_b200.tiles import TileDesc, TileMatrix


# ---- fake CUDA streams / events with vector clocks --------------------------------------------------
class FakeStream:
    _n = 0

    def __init__(self, *a, **k):
        FakeStream._n += 1
        self.name = f"s{FakeStream._n}"
        self.cuda_stream = 1000 + FakeStream._n
        self.clock = {}

    def tick(self):
        self.clock[self.name] = self.clock.get(self.name, 0) + 1
        return dict(self.clock)

    def _merge(self, other_clock):
        for k, v in other_clock.items():
            if v > self.clock.get(k, 0):
                self.clock[k] = v

    def wait_event(self, ev):
        assert ev.clock is not None, "waiting on an event that was never recorded"
        self._merge(ev.clock)

    def wait_stream(self, other):
        self._merge(other.clock)

    def synchronize(self):
        pass


class FakeEvent:
    def __init__(self, *a, **k):
        self.clock = None

    def record(self, stream=None):
        self.clock = dict((stream or _CUR[-1]).clock)


_CUR = []


@contextlib.contextmanager
def fake_stream_ctx(s):
    _CUR.append(s)
    try:
        yield
    finally:
        _CUR.pop()


class Probe(TiledCholesky):
    """One rank's schedule with recorded operations instead of kernels and collectives."""

    def __init__(self, A, lookahead, partition_tail=0):
        self.A, self.nt, self.b = A, A.nt, A.b
        self.grid, self.rank, self.lay = A.grid, A.rank, A.layout
        self.dev, self.world = A.device, A.grid.size
        self.cuda = True                      # take the CUDA branches of _run, with the fakes below
        self.group, self.lookahead, self.nslots, self.transport, self.tr = None, lookahead, 2, "nccl", None
        self.update_events = None
        self.tile_bytes = self.b * self.b * 8
        f64 = dict(dtype=torch.float64)
        self.work = torch.zeros(16, **f64)
        self.d_info = torch.zeros(1, dtype=torch.int32)
        self.panel = torch.zeros((2, max(self.nt - 1, 1), self.b, self.b), **f64) if self.world > 1 else None
        self.diag = torch.zeros((self.b, self.b), **f64) if self.world > 1 else None
        self._col_groups = [("col", q) for q in range(self.grid.Q)]
        self._build_plan()
        self.s_update, self.s_panel = FakeStream(), FakeStream()
        self.streams = {s.cuda_stream: s for s in (self.s_update, self.s_panel)}
        if partition_tail:
            # the SM partition of the tail (chol_partition_create): two more streams
            self.s_potrf, self.s_rest, self.tail_tasks = FakeStream(), FakeStream(), partition_tail
            self.streams.update({s.cuda_stream: s for s in (self.s_potrf, self.s_rest)})
        self.ops = []
        self.extra = {}                       # other tensors (residual scratch): data_ptr -> (name, tensor)

    # -- address -> symbolic region
    def region(self, ptr):
        tb = self.tile_bytes
        for name, t in (("A", self.A.buf), ("P", self.panel), ("D", self.diag), ("W", self.work)):
            if t is not None and t.data_ptr() <= ptr < t.data_ptr() + t.numel() * 8:
                return (name, (ptr - t.data_ptr()) // tb) if name in ("A", "P") else (name,)
        for base, (name, t) in self.extra.items():
            if base <= ptr < base + t.numel() * 8:
                return (name, (ptr - base) // tb)
        raise AssertionError(f"pointer {ptr:#x} outside every known buffer")

    def regions_of(self, t):
        first = self.region(t.data_ptr())
        if len(first) == 1:
            return {first}
        n = max(1, (t.numel() * 8 + self.tile_bytes - 1) // self.tile_bytes)
        return {(first[0], first[1] + i) for i in range(n)}

    def op(self, what, st, reads, writes):
        s = self.streams[st] if isinstance(st, int) else st
        assert _CUR and _CUR[-1] is s, f"{what}: launched on a stream that is not the current one"
        self.ops.append((what, s.name, s.tick(), set(reads), set(writes)))

    # -- recorded kernel entry points
    def _potrf_workspace(self, b):
        return 8

    def _k_potrf(self, a_ptr, info_base, st):
        self.op("potrf", st, {self.region(a_ptr)}, {self.region(a_ptr), ("W",), ("info",)})

    def _k_trsm_panel(self, l_ptr, work_ptr, tiles_ptr, ntiles, st):
        ptrs = (C.c_int64 * ntiles).from_address(tiles_ptr)
        tl = {self.region(p) for p in ptrs}
        self.op("trsm", st, tl | {self.region(l_ptr), ("W",)}, tl)

    def _k_update(self, tasks_ptr, ntasks, st, thin=False):
        rec = np.ctypeslib.as_array((C.c_int64 * (4 * ntasks)).from_address(tasks_ptr)).reshape(ntasks, 4)
        w = {self.region(c) for c in rec[:, 0].tolist()}
        r = {self.region(p) for p in rec[:, 1].tolist()} | {self.region(p) for p in rec[:, 2].tolist()}
        assert len(w) == ntasks, "two tasks of one launch write the same tile"
        self.op("update", st, r | w, w)

    def _k_tril(self, src_ptr, dst_ptr, st):
        self.op("tril", st, {self.region(src_ptr)}, {self.region(dst_ptr)})

    def _bcast(self, t, src, group):
        if isinstance(group, tuple):                      # column group: src is the process-row index
            mine = src == self.lay.p
        else:
            mine = src == self.rank
        regs = self.regions_of(t)
        self.op("bcast-send" if mine else "bcast-recv", _CUR[-1], regs if mine else set(), set() if mine else regs)


def happens_before(a, b):
    """a, b = (what, stream, clock, reads, writes); a was enqueued first."""
    return b[2].get(a[1], 0) >= a[2][a[1]]


def check_no_races(ops):
    last_w, readers = {}, {}
    bad = []
    for n, o in enumerate(ops):
        what, s, clk, reads, writes = o
        for reg in reads | writes:
            w = last_w.get(reg)
            if w is not None and not happens_before(ops[w], o):
                bad.append((ops[w][0], what, reg, "RAW/WAW"))
        for reg in writes:
            for r in readers.get(reg, ()):
                if r != n and not happens_before(ops[r], o):
                    bad.append((ops[r][0], what, reg, "WAR"))
            last_w[reg] = n
            readers[reg] = []
        for reg in reads - writes:
            readers.setdefault(reg, []).append(n)
    return bad


@pytest.fixture
def fake_cuda(monkeypatch):
    cur = FakeStream()
    _CUR.clear()
    _CUR.append(cur)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "stream", fake_stream_ctx)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _CUR[-1])
    yield cur
    _CUR.clear()


GRIDS = [(1, 1, 0), (1, 2, 0), (1, 2, 1), (2, 2, 0), (2, 2, 3), (2, 4, 0), (2, 4, 5), (2, 4, 7), (3, 2, 4)]


@pytest.mark.parametrize("P,Q,rank", GRIDS)
@pytest.mark.parametrize("lookahead", [True, False])
def test_factor_schedule_has_no_races(fake_cuda, P, Q, rank, lookahead):
    N, b = 16 * 11, 16
    M = TileMatrix(TileDesc(b, b, b * b, N, N, 0, 0, N, N, P, Q), rank, "cpu")
    pr = Probe(M, lookahead)
    pr._run(pr.d_tasks.data_ptr(), factor=True)
    kinds = {o[0] for o in pr.ops}
    assert "update" in kinds and ("potrf" in kinds or pr.world > 1)
    assert check_no_races(pr.ops) == []
    if lookahead and P * Q == 1:
        # the point of the lookahead: POTRF(k+1) is NOT ordered after the bulk of update k
        upd = [o for o in pr.ops if o[0] == "update"]
        pot = [o for o in pr.ops if o[0] == "potrf"]
        assert not happens_before(upd[2], pot[1]) and happens_before(upd[0], pot[1])


@pytest.mark.parametrize("P,Q,rank", GRIDS)
@pytest.mark.parametrize("tail", [3, 10, 1000])
def test_partitioned_tail_schedule_has_no_races(fake_cuda, P, Q, rank, tail):
    """With the SM partition the last steps' updates run on the `rest` stream and their POTRFs on the panel
    group's stream (tail=1000: from step 0 on): same dependences, two more streams."""
    N, b = 16 * 11, 16
    M = TileMatrix(TileDesc(b, b, b * b, N, N, 0, 0, N, N, P, Q), rank, "cpu")
    pr = Probe(M, True, partition_tail=tail)
    pr._run(pr.d_tasks.data_ptr(), factor=True)
    assert check_no_races(pr.ops) == []
    used = {o[1] for o in pr.ops}
    assert pr.s_rest.name in used or pr._tail_start() >= pr.nt
    if P * Q == 1:
        pot = [o for o in pr.ops if o[0] == "potrf"]
        assert any(o[1] == pr.s_potrf.name for o in pot) and pot[0][1] == pr.s_panel.name
        # still a lookahead: a partition POTRF is not ordered after the bulk update it overlaps
        k = next(i for i, o in enumerate(pot) if o[1] == pr.s_potrf.name)
        bulk = [o for o in pr.ops if o[0] == "update" and o[1] == pr.s_rest.name]
        assert any(not happens_before(u, pot[k]) for u in bulk)


@pytest.mark.parametrize("P,Q,rank", GRIDS)
def test_residual_schedule_has_no_races(fake_cuda, monkeypatch, P, Q, rank):
    N, b = 16 * 7, 16
    desc = TileDesc(b, b, b * b, N, N, 0, 0, N, N, P, Q)
    M, M0 = TileMatrix(desc, rank, "cpu"), TileMatrix(desc, rank, "cpu")
    M0.buf.zero_()
    pr = Probe(M, True)
    pr.extra[M0.buf.data_ptr()] = ("R", M0.buf)
    real_empty = torch.empty

    def tracking_empty(*a, **k):                      # the T_k scratch allocated inside residual()
        t = real_empty(*a, **k)
        if t.dtype == torch.float64 and t.dim() == 3 and t.shape[0] == 2:
            pr.extra[t.data_ptr()] = ("T", t)
        return t

    monkeypatch.setattr(torch, "empty", tracking_empty)
    monkeypatch.setattr(chol_mod, "_norms", lambda M_, ch: (1.0, 1.0))
    pr.residual(M0)
    assert any(o[0] == "tril" for o in pr.ops) or P * Q > 1
    assert check_no_races(pr.ops) == []


def test_the_checker_sees_a_missing_wait(fake_cuda):
    """Sanity of the checker itself: drop the event that orders TRSM(k+1) after the column update."""
    N, b = 16 * 6, 16
    M = TileMatrix(TileDesc.square(N, b), 0, "cpu")
    pr = Probe(M, True)
    real_wait = FakeStream.wait_event
    dropped = []

    def lossy_wait(self, ev):
        if self is pr.s_panel and len(dropped) < 40:      # the panel stream ignores its event waits
            dropped.append(ev)
            return
        real_wait(self, ev)

    FakeStream.wait_event = lossy_wait
    try:
        pr._run(pr.d_tasks.data_ptr(), factor=True)
    finally:
        FakeStream.wait_event = real_wait
    assert check_no_races(pr.ops) != []


@pytest.mark.parametrize("P,Q,rank", GRIDS)
@pytest.mark.parametrize("lookahead", [True, False])
def test_factor_from_host_schedule_has_no_races(fake_cuda, monkeypatch, P, Q, rank, lookahead):
    """The end-to-end entry adds an upload stream (tiles arrive group by group underneath step 0) and a
    download stream (each finished panel column leaves early): uploads are writes, downloads reads."""
    N, b = 16 * 9, 16
    M = TileMatrix(TileDesc(b, b, b * b, N, N, 0, 0, N, N, P, Q), rank, "cpu")
    pr = Probe(M, lookahead)
    hin, hout = torch.zeros_like(M.buf), torch.zeros_like(M.buf)
    real_copy = torch.Tensor.copy_
    lo, hi = M.buf.data_ptr(), M.buf.data_ptr() + M.buf.numel() * 8

    def recording_copy(dst, src, non_blocking=False):
        d_dev, s_dev = lo <= dst.data_ptr() < hi, lo <= src.data_ptr() < hi
        if d_dev or s_dev:
            pr.op("h2d" if d_dev else "d2h", _CUR[-1], pr.regions_of(src) if s_dev else set(),
                  pr.regions_of(dst) if d_dev else set())
            return dst
        return real_copy(dst, src, non_blocking)

    monkeypatch.setattr(torch.Tensor, "copy_", recording_copy)
    monkeypatch.setattr(torch.Tensor, "is_pinned", lambda self, *a, **k: True)
    pr.factor_from_host(hin, hout)
    pr.factor_from_host(hin, hout)                  # a second call reuses the streams: no race across calls either
    ups = set().union(*[o[4] for o in pr.ops if o[0] == "h2d"]) if any(o[0] == "h2d" for o in pr.ops) else set()
    downs = set().union(*[o[3] for o in pr.ops if o[0] == "d2h"]) if any(o[0] == "d2h" for o in pr.ops) else set()
    every = {("A", i) for i in range(M.layout.ntiles)}
    assert ups == every and downs == every          # each owned tile is uploaded once and downloaded once
    assert check_no_races(pr.ops) == []
    if lookahead and pr.step0_groups and len(pr.step0_groups) > 1:
        # the first update launches must not wait for the last upload group
        first_upd = next(o for o in pr.ops if o[0] == "update")
        last_up = [o for o in pr.ops if o[0] == "h2d"][-1]
        assert not happens_before(last_up, first_upd)
        if P * Q == 1:
            # lazy first steps: the panel chain gets S steps deep without waiting for the last upload group,
            # and step S (which touches every column) does wait for it
            S = pr.lazy_steps
            assert S >= 2
            n_up = sum(1 for o in pr.ops if o[0] == "h2d") // 2          # uploads of one call
            first_call = pr.ops[:next(i for i, o in enumerate(pr.ops) if o[0] == "h2d" and
                                      sum(1 for q in pr.ops[:i + 1] if q[0] == "h2d") > n_up)]
            pot = [o for o in first_call if o[0] == "potrf"]
            last_up1 = [o for o in first_call if o[0] == "h2d"][-1]
            assert not happens_before(last_up1, pot[S - 1])
            assert happens_before(last_up1, pot[S + 1])
