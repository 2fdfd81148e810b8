"""Panel transport between the ranks of a P x Q grid of B200s: copy-engine pushes + flag words.

The reference only carries the process grid in its descriptor (``p, q`` of CHAMELEON_Desc_Create,
v6_test.c:44-45) and leaves data movement to StarPU/MPI; the ArmoniK variant moves every tile
through the object store (worker_distrib.cpp:186,261).  Here the owner of a factored panel PUSHES it
over NVLink into receive slots of exactly the ranks whose trailing update reads it, and announces it
with a flag word; the readers wait with a stream memory operation, so no kernel is resident on a
waiting GPU (round 1's ncclBroadcast receivers spun on SMs and cost the update kernel 15 %).

What goes where (2D block-cyclic, tile (i,j) on rank (i % P, j % Q)):
  * L_kk (+ the inverted 128-blocks POTRF left in its workspace) goes from the diagonal owner to the
    other ranks of its process column that own tiles of panel k  ->  their ``diag`` slot;
  * panel tile (i,k) is read as the A operand by process row i % P and as the B operand by process
    column i % Q: the owner (row po) sends ALL its tiles to the ranks of its own row and only the
    tiles with i % Q == q to rank (p != po, q) — P + Q - 2 destinations per tile instead of P*Q - 1.
Slot reuse is guarded by credits: a reader raises its credit word on every rank when update k (panel)
or TRSM k (diag) is enqueued behind its last read; the sender's copy waits for the credit of the
step that used the slot before.  All words are cyclic 32-bit counters ``epoch * nt + k + 1`` so
nothing is ever reset.

Host logic only; the primitives (`_send`, `_wait`, `_post`, `_open`) are the C ABI's
chol_peer_* / chol_flag_* calls and are replaced by a shared-memory double in the CPU tests.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .grid import panel_slots

FLAG_BYTES = 4096          # flag words, padded so the data that follows stays page aligned
F_PANEL_DATA, F_DIAG_DATA, F_PANEL_CREDIT, F_DIAG_CREDIT = 0, 1, 2, 3   # blocks of `world` words


@dataclass
class Push:
    """One push to `peer`: `count` tiles from local address `src` to byte offset `dst_off` of the
    peer's buffer; credit / flag are word indices (or None)."""
    peer: int
    dst_off: int
    src: int
    tile_bytes: int
    count: int
    dst_stride: int = 1
    src_stride: int = 1
    credit: int | None = None
    credit_value: int = 0
    flag: int | None = None
    flag_value: int = 0


def progression(ts: list[int]) -> list[tuple[int, int, int]]:
    """Split a sorted index list into (first, count, stride) runs (one run when it is an
    arithmetic progression, which a block-cyclic subset always is)."""
    runs, n = [], len(ts)
    a = 0
    while a < n:
        if a + 1 == n:
            runs.append((ts[a], 1, 1))
            break
        step = ts[a + 1] - ts[a]
        e = a + 1
        while e + 1 < n and ts[e + 1] - ts[e] == step:
            e += 1
        runs.append((ts[a], e - a + 1, step))
        a = e + 1
    return runs


def panel_subset(nt: int, P: int, Q: int, k: int, po: int, reader: tuple[int, int]) -> list[int]:
    """Positions t (in the owner row's ascending list of tiles (i,k), i > k, i % P == po) that the
    rank at grid coordinates `reader` reads in update k."""
    p, q = reader
    lo = k + 1
    i0 = lo + ((po - lo) % P)
    rows = range(i0, nt, P)
    if p == po:
        return list(range(len(rows)))
    return [t for t, i in enumerate(rows) if i % Q == q]


class PeerTransport:
    NSLOTS = 4      # panel receive slots (lookahead 1 needs 2; the rest absorbs skew between ranks)
    NDIAG = 2

    def __init__(self, nt: int, b: int, grid, rank: int, work_doubles: int, group=None):
        self.nt, self.b, self.grid, self.rank, self.group = nt, b, grid, rank, group
        self.world = grid.size
        self.p, self.q = grid.coords(rank)
        self.tile_bytes = b * b * 8
        self.work_bytes = work_doubles * 8
        self.nslots = max(2, min(self.NSLOTS, max(nt - 1, 2)))
        self.diag_stride = -(-(self.tile_bytes + self.work_bytes) // 256) * 256
        self.slot_tiles = max(nt - 1, 1)
        self.off_diag = FLAG_BYTES
        self.off_panel = self.off_diag + self.NDIAG * self.diag_stride
        self.nbytes = self.off_panel + self.nslots * self.slot_tiles * self.tile_bytes
        assert 4 * self.world * 4 <= FLAG_BYTES
        self._plan_cache, self._diag_cache = {}, {}
        self.epoch = 0
        self.base = 0
        self.local = 0            # address of this rank's buffer
        self.peer_base = {}       # rank -> address of its buffer as mapped here
        self._open()

    # ---- addresses -------------------------------------------------------------------------------
    def flag_index(self, block: int, who: int) -> int:
        return block * self.world + who

    def panel_slot_ptr(self, k: int) -> int:
        return self.local + self.off_panel + (k % self.nslots) * self.slot_tiles * self.tile_bytes

    def diag_slot(self, k: int) -> int:
        return (k // self.grid.Q) % self.NDIAG

    def diag_tile_ptr(self, k: int) -> int:
        return self.local + self.off_diag + self.diag_slot(k) * self.diag_stride

    def diag_work_ptr(self, k: int) -> int:
        return self.diag_tile_ptr(k) + self.tile_bytes

    # ---- per-run state ---------------------------------------------------------------------------
    def begin_run(self) -> None:
        """Every rank calls this once per pass of the schedule (factor, residual, ...), collectively."""
        self.base = (self.epoch * self.nt) & 0xFFFFFFFF
        self.epoch += 1

    def _val(self, k: int) -> int:
        return (self.base + k + 1) & 0xFFFFFFFF

    def end_run(self, stream) -> None:
        """After the streams of the pass have been joined into `stream`: release every slot."""
        self._post_credit(F_PANEL_CREDIT, self.nt - 1, stream)
        if self.grid.P > 1:
            self._post_credit(F_DIAG_CREDIT, self.nt - 1, stream)

    # ---- L_kk down the process column --------------------------------------------------------------
    def diag_readers(self, k: int) -> list[int]:
        """Ranks (other than the diagonal owner) of process column k % Q that own tiles of panel k."""
        if k not in self._diag_cache:
            g = self.grid
            _, groups = panel_slots(self.nt, g.P, k)
            self._diag_cache[k] = [g.rank_of(p, k % g.Q) for p, _, cnt in groups if cnt and p != k % g.P]
        return self._diag_cache[k]

    def send_diag(self, k: int, l_ptr: int, work_ptr: int, ready_stream, send_stream_of) -> None:
        Q, tb = self.grid.Q, self.tile_bytes
        prev = k - self.NDIAG * Q          # the step that used this diag slot before
        cval = self._val(prev) if prev >= 0 else self.base
        pushes = []
        for r in self.diag_readers(k):
            off = self.off_diag + self.diag_slot(k) * self.diag_stride
            pushes.append(Push(r, off, l_ptr, tb, 1, credit=self.flag_index(F_DIAG_CREDIT, r), credit_value=cval))
            pushes.append(Push(r, off + tb, work_ptr, self.work_bytes, 1,
                               flag=self.flag_index(F_DIAG_DATA, self.rank), flag_value=self._val(k)))
        if pushes:
            self._send(pushes, ready_stream, send_stream_of)

    def wait_diag(self, k: int, stream) -> None:
        g = self.grid
        owner = g.rank_of(k % g.P, k % g.Q)
        if self.rank in self.diag_readers(k):
            self._wait(self.flag_index(F_DIAG_DATA, owner), self._val(k), stream)

    def release_diag(self, k: int, stream) -> None:
        """TRSM k has been enqueued on `stream` (or this rank had none): the diag slot may be reused."""
        if self.grid.P > 1:
            self._post_credit(F_DIAG_CREDIT, k, stream, column_only=True)

    # ---- the factored panel ------------------------------------------------------------------------
    def panel_plan(self, k: int):
        """[(owner rank, po, first_slot, {reader rank: [(t0, count, stride)]})] for panel k."""
        if k in self._plan_cache:
            return self._plan_cache[k]
        g, nt = self.grid, self.nt
        _, groups = panel_slots(nt, g.P, k)
        out = self._plan_cache[k] = []
        for po, first, cnt in groups:
            if not cnt:
                continue
            owner = g.rank_of(po, k % g.Q)
            dests = {}
            for r in range(self.world):
                if r == owner:
                    continue
                ts = panel_subset(nt, g.P, g.Q, k, po, g.coords(r))
                if ts:
                    dests[r] = progression(ts)
            out.append((owner, po, first, dests))
        return out

    def send_panel(self, k: int, src_ptr: int, ready_stream, send_stream_of) -> None:
        """Owner side: `src_ptr` = address of this rank's first tile of panel k (its tiles of the
        panel are contiguous, ascending i)."""
        tb = self.tile_bytes
        prev = k - self.nslots
        cval = self._val(prev) if prev >= 0 else self.base
        pushes = []
        for owner, _, first, dests in self.panel_plan(k):
            if owner != self.rank:
                continue
            for r, runs in dests.items():
                base_off = self.off_panel + ((k % self.nslots) * self.slot_tiles + first) * tb
                for n, (t0, cnt, stride) in enumerate(runs):
                    last = n + 1 == len(runs)
                    pushes.append(Push(r, base_off + t0 * tb, src_ptr + t0 * tb, tb, cnt, stride, stride,
                                       credit=self.flag_index(F_PANEL_CREDIT, r) if n == 0 else None, credit_value=cval,
                                       flag=self.flag_index(F_PANEL_DATA, self.rank) if last else None,
                                       flag_value=self._val(k)))
        if pushes:
            self._send(pushes, ready_stream, send_stream_of)

    def wait_panel(self, k: int, stream) -> None:
        for owner, _, _, dests in self.panel_plan(k):
            if owner != self.rank and self.rank in dests:
                self._wait(self.flag_index(F_PANEL_DATA, owner), self._val(k), stream)

    def release_panel(self, k: int, stream) -> None:
        """Update k has been enqueued on `stream`: panel slot k % nslots may be overwritten."""
        self._post_credit(F_PANEL_CREDIT, k, stream)

    def _post_credit(self, block: int, k: int, stream, column_only: bool = False) -> None:
        g = self.grid
        peers = [r for r in range(self.world) if r != self.rank and (not column_only or r % g.Q == self.q)]
        if peers:
            self._post([(r, self.flag_index(block, self.rank)) for r in peers], self._val(k), stream)

    # ---- primitives: the C ABI (replaced by a shared-memory double in tests/) -------------------------
    def _open(self) -> None:
        import torch.distributed as dist
        out = C.c_void_p()
        _lib.call("chol_peer_alloc", self.nbytes, C.byref(out))
        self.local = out.value
        handle = (C.c_ubyte * 64)()
        _lib.call("chol_peer_export", self.local, handle)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        for r, h in enumerate(handles):
            if r == self.rank:
                continue
            buf = (C.c_ubyte * 64).from_buffer_copy(h)
            _lib.call("chol_peer_open", buf, C.byref(out))
            self.peer_base[r] = out.value
        dist.barrier(group=self.group)

    def close(self) -> None:
        """Unmap the peers' buffers and free ours (collective: nobody may still be pushing)."""
        if not self.local:
            return
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier(group=self.group)
        for ptr in self.peer_base.values():
            _lib.call("chol_peer_close", ptr)
        self.peer_base = {}
        if dist.is_initialized():
            dist.barrier(group=self.group)
        _lib.call("chol_peer_free", self.local)
        self.local = 0

    def _send(self, pushes: list[Push], ready_stream, send_stream_of) -> None:
        arr = (_lib.Xfer * len(pushes))()
        for x, p in zip(arr, pushes):
            base = self.peer_base[p.peer]
            x.dst, x.src, x.tile_bytes, x.count = base + p.dst_off, p.src, p.tile_bytes, p.count
            x.dst_stride, x.src_stride = p.dst_stride, p.src_stride
            x.credit = self.local + 4 * p.credit if p.credit is not None else None
            x.credit_value = p.credit_value
            x.flag = base + 4 * p.flag if p.flag is not None else None
            x.flag_value = p.flag_value
            x.stream = send_stream_of(p.peer)
        _lib.call("chol_peer_send", arr, len(pushes), ready_stream)

    def _wait(self, flag: int, value: int, stream) -> None:
        _lib.call("chol_flag_wait", self.local + 4 * flag, value, stream)

    def _post(self, targets: list[tuple[int, int]], value: int, stream) -> None:
        ptrs = (C.c_void_p * len(targets))(*[self.peer_base[r] + 4 * f for r, f in targets])
        _lib.call("chol_flag_post", ptrs, len(targets), value, stream)

    # ---- bookkeeping for reports ---------------------------------------------------------------------
    def bytes_received_per_run(self) -> int:
        """Bytes other ranks push into this rank's slots during one factorization."""
        n = 0
        for k in range(self.nt - 1):
            for owner, _, _, dests in self.panel_plan(k):
                if self.rank in dests:
                    n += sum(c for _, c, _ in dests[self.rank]) * self.tile_bytes
            if self.rank in self.diag_readers(k):
                n += self.tile_bytes + self.work_bytes
        return n


def as_np(ptr: int, nbytes: int) -> np.ndarray:
    return np.ctypeslib.as_array((C.c_ubyte * nbytes).from_address(ptr))
