"""Right-looking tiled Cholesky on one or several B200s: the host side.

Replaces ``CHAMELEON_dpotrf_Tile(ChamLower, descA)`` (v6_test.c:56) and the client's wave loop
(client_distrib.cpp v1:278-333 / v2:506-565): for each k, POTRF(k,k); TRSM(i,k) for i>k; then
SYRK(i,i) / GEMM(i,j) for i>=j>k.  Python builds that DAG once as a *plan* (per-step device
task lists of tile pointers) and then only enqueues kernels of ``libchol_b200.so`` through
ctypes; torch provides the buffers, streams, events and the NCCL broadcasts.

Schedule (per rank; one rank per GPU, tiles 2D block-cyclic, see grid.py):
  * panel stream (high priority): POTRF of the diagonal tile on its owner, L_kk and the inverted
    diagonal blocks pushed down the owner's process column, TRSM of the panel tiles on their
    owners (one grouped launch sequence), the factored panel pushed to the ranks that read it
    (transport.py: copy-engine peer copies + flag words, no kernel resident on a waiting GPU;
    CHOL_PANEL_TRANSPORT=nccl keeps round 1's ncclBroadcast path for A/B runs);
  * update stream: the fused SYRK+GEMM trailing update of step k as grouped launches over every
    local tile (i,j), i>=j>k, in three stages: the diagonal tile (k+1,k+1) — POTRF(k+1) starts
    right after it —, the rest of column k+1 — then TRSM(k+1) —, and everything else, which
    overlaps panel step k+1: lookahead of depth 1;
  * with host-resident input (factor_from_host) a third and fourth stream upload the tiles in
    storage order underneath step 0 and copy each finished panel column back.
Nothing synchronises with the host between steps; LAPACK ``info`` is a device int read at the end.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .grid import panel_slots
from .tiles import TileMatrix


class TiledCholesky:
    """Plan + executor for the in-place factorization of a TileMatrix (lower, A = L L^T)."""
    trace = None
    tr = None
    thin_tasks = 0
    _potrf_on_partition = False
    lazy_steps = 1
    lazy_index = None
    d_lazy = None
    s_potrf = None      # streams of the SM partition (chol_partition_create), None = no partition
    s_rest = None
    tail_tasks = 0
    _part = None
    _su = None          # the update stream of the step being enqueued (s_update or s_rest)

    def __init__(self, A: TileMatrix, group=None, lookahead: bool = True):
        self.A = A
        self.nt, self.b = A.nt, A.b
        self.grid, self.rank, self.lay = A.grid, A.rank, A.layout
        self.dev = A.device
        self.cuda = self.dev.type == "cuda"
        self.world = self.grid.size
        self.group = group
        self.lookahead = lookahead
        if self.world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                raise RuntimeError("a P x Q grid with more than one rank needs torch.distributed initialised")
            if dist.get_world_size(group) != self.world:
                raise RuntimeError(f"process group has {dist.get_world_size(group)} ranks, grid needs {self.world}")
        b, nt = self.b, self.nt
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.tile_bytes = b * b * 8
        # workspace: inverted 128x128 diagonal blocks of L_kk (chol_potrf_tile -> chol_trsm_tiles)
        self.work = torch.empty(max(self._potrf_workspace(b) // 8, 1), **f64)
        self.d_info = torch.zeros(1, dtype=torch.int32, device=self.dev)
        # receive buffers (only with more than one rank): panel column k, two slots for lookahead
        self.panel = None
        self.diag = None
        # How the factored panel reaches the other ranks: "peer" = pushes into IPC-mapped receive
        # slots on the copy engines + flag words (transport.py, default on CUDA); "nccl" = one
        # ncclBroadcast per owner row (round 1; also what the gloo CPU tests exercise).
        default = "peer" if self.cuda else "nccl"
        self.transport = os.environ.get("CHOL_PANEL_TRANSPORT", default) if self.world > 1 else "nccl"
        if self.transport not in ("nccl", "peer"):
            raise ValueError("CHOL_PANEL_TRANSPORT must be 'nccl' or 'peer'")
        self.nslots = 2
        self.tr = None
        if self.world > 1:
            if self.transport == "peer":
                self.tr = self._make_transport()
                self.nslots = self.tr.nslots
            else:
                self.panel = torch.empty((2, max(nt - 1, 1), b, b), **f64)
                self.diag = torch.empty((b, b), **f64)
        self._col_groups = None
        if self.world > 1 and self.grid.P > 1:
            # dist.new_group is collective over ALL ranks: create every column group here, in the
            # same order everywhere, never lazily inside the factorization
            self._make_column_groups()
        # bulk updates of at most this many tasks run at 1 CTA/SM (see _run); 0 = never.  Measured on 8 B200
        # (N=65536): 0 -> 393 ms, 40 -> 386, 80 -> 388, 160 -> 414; on one GPU it makes no difference, so it is
        # on only when several ranks share the panel chain.  CHOL_THIN_TASKS overrides.
        self.thin_tasks = int(os.environ.get("CHOL_THIN_TASKS", "48" if self.world > 1 else "0"))
        self.update_events = None   # set to [] to time every trailing-update launch (bench.py roofline)
        self.trace = None           # set to [] to record a CUDA-event timeline of the next pass (tools/)
        self._build_plan()
        if self.cuda:
            self.s_update = torch.cuda.Stream(self.dev)
            self.s_panel = torch.cuda.Stream(self.dev, priority=-1)
            self._make_partition()
            if self.tr is not None:
                # one send stream per peer (copies to different peers run on different copy engines)
                # and one for the credit words, so neither ever sits in front of a kernel
                self.s_sends = {r: torch.cuda.Stream(self.dev, priority=-1) for r in range(self.world) if r != self.rank}
                self.s_credit = torch.cuda.Stream(self.dev, priority=-1)
        else:
            self.s_update = self.s_panel = None
            self.s_sends, self.s_credit = {}, None

    def _make_partition(self) -> None:
        """SM partition for the tail of the factorization (include/chol_b200.h, chol_partition_create):
        once the bulk of a step's update has at most `tail_tasks` tile tasks, the panel chain bounds the
        step; from then on the updates run on the `rest` group of SMs and POTRF on its private group, so
        none of its ~24 short dependent kernels waits for an update CTA to retire (measured on B200: a
        POTRF tile takes 0.6 ms alone, 1.3-1.5 ms under a running update, 0.74 ms on its own 16 SMs;
        N=16384: 51.9 -> 51.1 ms, 8 GPUs N=65536: 393.5 -> 390.1 ms).
        OPT-IN (CHOL_PANEL_SMS=16): a process that uses green contexts is killed by Nsight Compute 2025.2
        (observed: exit status 9 at the first launch on a green-context stream), and a 1 % gain does not
        justify a default that cannot be profiled."""
        sms = int(os.environ.get("CHOL_PANEL_SMS", "0"))
        self.tail_tasks = int(os.environ.get("CHOL_TAIL_TASKS", "28" if self.world == 1 else "48"))
        if sms <= 0 or self.tail_tasks <= 0 or self.nt < 3:
            return
        key = self.dev.index if self.dev.index is not None else torch.cuda.current_device()
        if key not in _partitions:
            part = _lib.Partition()
            try:
                _lib.call("chol_partition_create", key, sms, C.byref(part))
                _partitions[key] = (part, torch.cuda.ExternalStream(part.panel_stream, device=self.dev),
                                    torch.cuda.ExternalStream(part.rest_stream, device=self.dev))
            except _lib.CholError as e:
                if os.environ.get("CHOL_PANEL_SMS"):
                    raise
                import sys
                sys.stderr.write(f"[chol] no SM partition ({e}); using the two-stream schedule\n")
                _partitions[key] = None
        if _partitions[key] is not None:
            self._part, self.s_potrf, self.s_rest = _partitions[key]
            self.thin_tasks = 0           # the partition replaces the 1-CTA/SM tail updates

    def _tail_start(self) -> int:
        """First step from which every step's bulk update (part b) has at most tail_tasks tasks."""
        k = self.nt
        while k > 0 and self.step_tasks[k - 1][3] - self.step_tasks[k - 1][2] <= self.tail_tasks:
            k -= 1
        return k

    def _make_transport(self):
        """The peer-push transport of this geometry (tests substitute a shared-memory double)."""
        from .transport import PeerTransport
        return PeerTransport(self.nt, self.b, self.grid, self.rank, self.work.numel(), self.group)

    def _send_stream_of(self, r: int) -> int:
        return self.s_sends[r].cuda_stream if self.s_sends else 0

    def close(self) -> None:
        """Release the transport's IPC mappings (collective over the ranks).  Optional: process exit
        does the same."""
        if self.tr is not None:
            self.tr.close()
            self.tr = None

    # ---- kernel entry points (the C ABI); tests override these on CPU tensors ----------------
    def _potrf_workspace(self, b: int) -> int:
        return _lib.load().chol_potrf_tile_workspace(b)

    def _k_potrf(self, a_ptr: int, info_base: int, st: int) -> None:
        _lib.call("chol_potrf_tile", self.b, a_ptr, self.b, self.work.data_ptr(), self.d_info.data_ptr(), info_base, st)

    def _k_trsm_panel(self, l_ptr: int, work_ptr: int, tiles_ptr: int, ntiles: int, st: int) -> None:
        _lib.call("chol_trsm_tiles", self.b, l_ptr, self.b, work_ptr, tiles_ptr, ntiles, self.b, None, st)

    def _k_update(self, tasks_ptr: int, ntasks: int, st: int, thin: bool = False) -> None:
        b = self.b
        occ = 1 if thin else 2
        if self.update_events is None:
            _lib.call("chol_gemm_tasks_ex", tasks_ptr, ntasks, b, b, b, b, b, b, -1.0, 1.0, occ, st)
            return
        # CUDA events on the launching stream around this one launch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self._su or self.s_update)
        _lib.call("chol_gemm_tasks_ex", tasks_ptr, ntasks, b, b, b, b, b, b, -1.0, 1.0, occ, st)
        e1.record(self._su or self.s_update)
        first = (tasks_ptr - self.d_tasks.data_ptr()) // 32
        nsyrk = int(self.tasks_host[first:first + ntasks, 3].sum()) if 0 <= first < len(self.tasks_host) else 0
        self.update_events.append((e0, e1, (2 * ntasks - nsyrk) * float(b) ** 3))

    def _k_tril(self, src_ptr: int, dst_ptr: int, st: int) -> None:
        _lib.call("chol_tile_tril", self.b, src_ptr, self.b, dst_ptr, self.b, st)

    def _bcast(self, t: torch.Tensor, src: int, group) -> None:
        import torch.distributed as dist
        dist.broadcast(t, src=dist.get_global_rank(group, src) if group is not None else src, group=group)

    # ---- plan ----------------------------------------------------------------------------------
    def _panel_ptrs(self, k: int, base_local: int) -> np.ndarray:
        """Address, on this rank, of every tile (i, k): local storage if owned, else the receive
        buffer slot of step k.  Entries for i <= k are 0."""
        nt, tb = self.nt, self.tile_bytes
        ptr = np.zeros(nt, dtype=np.int64)
        if self.world == 1:
            rows = np.arange(k + 1, nt)
            ptr[rows] = base_local + (self.lay.col_start[k] + rows - k) * tb
            return ptr
        slot, _ = panel_slots(nt, self.grid.P, k)
        if self.tr is not None:
            pbase = self.tr.panel_slot_ptr(k)
        else:
            pbase = self.panel.data_ptr() + (k % self.nslots) * self.panel.stride(0) * 8
        mine_col = (k % self.grid.Q) == self.lay.q
        for i in range(k + 1, nt):
            if mine_col and i % self.grid.P == self.lay.p:
                ptr[i] = base_local + self.lay.index(i, k) * tb
            else:
                ptr[i] = pbase + slot[i] * tb
        return ptr

    def _update_tasks(self, k: int, c_base: int, a_ptr: np.ndarray) -> tuple[np.ndarray, int, int]:
        """Task records (C, A, B, flags) of the trailing update of step k for the local tiles
        (i,j), i>=j>k: C_ij -= A_ik A_jk^T, lower only when i == j (client loop v1:307-329).
        Column k+1 first (part a), and inside it the diagonal tile (k+1,k+1) first when owned;
        returns (tasks[n,4], n_diag in {0,1}, n_part_a)."""
        lay, tb, P = self.lay, self.tile_bytes, self.grid.P
        cols = [j for j in lay.cols if j > k]
        parts_a, parts_b = [], []
        for j in cols:
            rows = np.asarray(lay.rows_in_col(j, j - 1), dtype=np.int64)  # owned i >= j
            if rows.size == 0:
                continue
            cidx = lay.col_start[j] + (rows - lay.col_first_row[j]) // P
            rec = np.empty((rows.size, 4), dtype=np.int64)
            rec[:, 0] = c_base + cidx * tb
            rec[:, 1] = a_ptr[rows]
            rec[:, 2] = a_ptr[j]
            rec[:, 3] = (rows == j).astype(np.int64)
            (parts_a if j == k + 1 else parts_b).append((rows, rec))
        na = sum(r.shape[0] for _, r in parts_a)
        nd = 1 if (parts_a and int(parts_a[0][0][0]) == k + 1) else 0
        if parts_b:
            rows_b = np.concatenate([r for r, _ in parts_b])
            rec_b = np.concatenate([r for _, r in parts_b])
            if k > 0:
                # row-major over (i, j): consecutive tasks share the A operand (tile (i,k)) in L2
                rec_b = rec_b[np.argsort(rows_b, kind="stable")]
            # step 0 stays column-major = storage order (consecutive tasks share the B operand):
            # factor_from_host uploads the tiles in that order and releases the update group by group
            recs = [r for _, r in parts_a] + [rec_b]
        else:
            recs = [r for _, r in parts_a]
        if not recs:
            return np.zeros((0, 4), dtype=np.int64), 0, 0
        return np.concatenate(recs), nd, na

    def _build_plan(self) -> None:
        nt, tb = self.nt, self.tile_bytes
        base = self.A.buf.data_ptr()
        all_tasks, self.step_tasks = [], []   # step_tasks[k] = (offset, n_diag, n_a, n_total)
        trsm_ptrs, self.step_trsm = [], []    # step_trsm[k] = (offset, count) of owned panel tiles
        off = toff = 0
        for k in range(nt):
            a_ptr = self._panel_ptrs(k, base)
            rec, nd, na = self._update_tasks(k, base, a_ptr)
            all_tasks.append(rec)
            self.step_tasks.append((off, nd, na, rec.shape[0]))
            off += rec.shape[0]
            if (k % self.grid.Q) == self.lay.q:
                rows = np.asarray(self.lay.rows_in_col(k, k), dtype=np.int64)
            else:
                rows = np.zeros(0, dtype=np.int64)
            trsm_ptrs.append(a_ptr[rows])
            self.step_trsm.append((toff, rows.size))
            toff += rows.size
        tasks = np.concatenate(all_tasks) if off else np.zeros((1, 4), dtype=np.int64)
        tptr = np.concatenate(trsm_ptrs) if toff else np.zeros(1, dtype=np.int64)
        self.tasks_host = tasks
        self.d_tasks = torch.from_numpy(tasks).to(self.dev)
        self.d_trsm_ptrs = torch.from_numpy(tptr).to(self.dev)
        self.n_update_tasks = off
        self.step0_head, self.step0_groups = self._step0_groups()
        self._build_lazy_plan()

    def _step0_groups(self, ngroups: int = 8):
        """Upload/compute pipeline of factor_from_host for step 0.  Returns (head_hi, groups): local
        tiles [0, head_hi) are columns 0 and 1 (panel 0 and part a); every later local tile is the C
        operand of exactly one part-b task of step 0, in the same (column-major) order, so group g =
        (task0, task1, tile_lo, tile_hi) can start as soon as tiles [0, tile_hi) are on the device."""
        lay = self.lay
        cols = [j for j in lay.cols if j >= 2]
        head_hi = lay.col_start[cols[0]] if cols else lay.ntiles
        groups = []
        total = lay.ntiles - head_hi
        if total > 0:
            target = max(1, -(-total // ngroups))
            t0, lo = 0, head_hi
            for n, j in enumerate(cols):
                nxt = lay.col_start[cols[n + 1]] if n + 1 < len(cols) else lay.ntiles
                if nxt - lo >= target or n + 1 == len(cols):
                    if nxt > lo:
                        groups.append((t0, t0 + (nxt - lo), lo, nxt))
                    t0 += nxt - lo
                    lo = nxt
        return head_hi, groups

    def _build_lazy_plan(self, max_steps: int = 4) -> None:
        """Host-resident input on one rank (factor_from_host): the upload of a large matrix takes longer than
        step 0's update (N=65536 over PCIe: 0.32 s against 0.13 s), and a right-looking step touches every
        column, so step 1 would wait for the last tile.  The first S steps are therefore applied LAZILY: while
        the later column groups are still arriving, steps 0..S-1 run on the columns that are already there
        (the head and upload group 0, which hold every column <= S: all their panels and the first panel after
        them); each later group receives its updates 0..S-1, in that order, as soon as it has arrived; from
        step S on everything is as usual.  Every tile still sees its updates in the same order, one launch per
        step, so the factor is bit-identical.  Here: the part-b tasks of steps 1..S-1 cut by upload group."""
        self.lazy_steps, self.lazy_index, self.d_lazy = 1, {}, None
        if self.world != 1 or len(self.step0_groups) < 2:
            return
        lay, tb, base = self.lay, self.tile_bytes, self.A.buf.data_ptr()
        hi0 = self.step0_groups[0][3]                      # tiles [0, hi0): head + group 0
        last_col = max((j for j in lay.cols if lay.col_start[j] < hi0 and
                        (lay.col_start[j] + len(lay.rows_in_col(j, j - 1))) <= hi0), default=0)
        S = min(max_steps, last_col, self.nt - 2)
        if S < 2:
            return
        bounds = [(g[2], g[3]) for g in self.step0_groups]
        bounds[0] = (self.step0_head, bounds[0][1])
        recs, off = [], 0
        for k in range(1, S):
            o, nd, na, ntot = self.step_tasks[k]
            part_b = self.tasks_host[o + na:o + ntot]
            tile = (part_b[:, 0] - base) // tb
            for g, (lo, hi) in enumerate(bounds):
                sel = part_b[(tile >= lo) & (tile < hi)]
                self.lazy_index[(k, g)] = (off, sel.shape[0])
                recs.append(sel)
                off += sel.shape[0]
            assert sum(self.lazy_index[(k, g)][1] for g in range(len(bounds))) == ntot - na
        self.lazy_steps = S
        self.d_lazy = torch.from_numpy(np.concatenate(recs) if off else np.zeros((1, 4), dtype=np.int64)).to(self.dev)

    def _make_column_groups(self) -> None:
        import torch.distributed as dist
        base_ranks = list(range(self.world)) if self.group is None else dist.get_process_group_ranks(self.group)
        opts = None
        if self.cuda:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        self._col_groups = []
        for qq in range(self.grid.Q):
            members = [base_ranks[self.grid.rank_of(p, qq)] for p in range(self.grid.P)]
            self._col_groups.append(dist.new_group(members, pg_options=opts) if opts else dist.new_group(members))
        # touch every communicator once now (NCCL may set them up lazily, which blocks the host —
        # that must not happen in the middle of the asynchronous factorization)
        tok = torch.zeros(1, dtype=torch.float64, device=self.dev)
        dist.broadcast(tok, src=base_ranks[0], group=self.group)
        dist.broadcast(tok, src=base_ranks[self.grid.rank_of(0, self.lay.q)], group=self._col_groups[self.lay.q])
        if self.cuda:
            torch.cuda.synchronize(self.dev)

    def _column_group(self, q: int):
        """Process group of the P ranks of grid column q (for the L_kk broadcast)."""
        return None if self.grid.P == 1 else self._col_groups[q]

    # ---- execution -------------------------------------------------------------------------------
    def _stream_ptr(self, s) -> int:
        return s.cuda_stream if s is not None else 0

    def _panel_potrf(self, k: int, factor: bool = True) -> None:
        """First half of panel step k (panel stream): POTRF of tile (k,k) on its owner and the
        broadcast of L_kk + its inverted diagonal blocks down the owner's process column.  Needs
        only tile (k,k) up to date, so it starts as soon as the diagonal SYRK of step k-1 is done."""
        if not factor:
            return
        g, lay, nt = self.grid, self.lay, self.nt
        st = self._stream_ptr(self.s_panel)
        kq, kp = k % g.Q, k % g.P
        in_col = kq == lay.q
        is_diag = in_col and kp == lay.p
        if is_diag and self._potrf_on_partition:
            # POTRF on its private SMs: it keeps its place in the panel stream's order (everything enqueued
            # there so far comes first, everything after it waits for it), only its kernels run elsewhere
            sp = self.s_potrf
            ev = torch.cuda.Event()
            ev.record(self.s_panel)
            sp.wait_event(ev)
            with torch.cuda.stream(sp):
                self._mark("potrf0", k, sp)
                self._k_potrf(self.A.tile_ptr(k, k), k * self.b, self._stream_ptr(sp))
                self._mark("potrf1", k, sp)
                ev = torch.cuda.Event()
                ev.record(sp)
            self.s_panel.wait_event(ev)
        elif is_diag:
            self._mark("potrf0", k, self.s_panel)
            self._k_potrf(self.A.tile_ptr(k, k), k * self.b, st)
            self._mark("potrf1", k, self.s_panel)
        self._l_ptr = self.A.tile_ptr(k, k) if is_diag else 0
        self._w_ptr = self.work.data_ptr()
        if g.P > 1 and in_col and k + 1 < nt:
            if self.tr is not None:
                if is_diag:
                    self.tr.send_diag(k, self._l_ptr, self._w_ptr, st, self._send_stream_of)
                elif self.rank in self.tr.diag_readers(k):
                    self.tr.wait_diag(k, st)
                    self._mark("diag_in", k, self.s_panel)
                    self._l_ptr, self._w_ptr = self.tr.diag_tile_ptr(k), self.tr.diag_work_ptr(k)
                return
            cg = self._column_group(kq)
            ltile = self.A.tile(k, k) if is_diag else self.diag
            self._bcast(ltile, kp, cg)
            self._bcast(self.work, kp, cg)
            self._l_ptr = ltile.data_ptr()

    def _panel_rest(self, k: int, factor: bool = True) -> None:
        """Second half of panel step k: TRSM of the owned panel tiles (needs column k up to date)
        and the broadcast of the factored panel to every rank.  With factor=False only the
        broadcasts run (residual mode)."""
        g, lay, nt = self.grid, self.lay, self.nt
        st = self._stream_ptr(self.s_panel)
        kq = k % g.Q
        if factor:
            toff, cnt = self.step_trsm[k]
            if cnt:
                self._mark("trsm0", k, self.s_panel)
                self._k_trsm_panel(self._l_ptr, self._w_ptr, self.d_trsm_ptrs.data_ptr() + toff * 8, cnt, st)
                self._mark("trsm1", k, self.s_panel)
        if self.world > 1 and k + 1 < nt:
            if self.tr is not None:
                if kq == lay.q:
                    rows = lay.rows_in_col(k, k)
                    if len(rows):
                        self.tr.send_panel(k, self.A.tile_ptr(rows[0], k), st, self._send_stream_of)
                self.tr.wait_panel(k, st)
                self._mark("panel_in", k, self.s_panel)
                return
            _, groups = panel_slots(nt, g.P, k)
            for p, first, cnt in groups:
                if cnt == 0:
                    continue
                root = g.rank_of(p, kq)
                if root == self.rank:
                    s0 = lay.index(lay.rows_in_col(k, k)[0], k)
                    buf = self.A.buf[s0:s0 + cnt]
                else:
                    buf = self.panel[k % self.nslots, first:first + cnt]
                self._bcast(buf, root, self.group)

    def _run(self, update_tasks_ptr: int, factor: bool, pre_update=None, post_panel=None, step0_gates=None) -> None:
        nt = self.nt
        cuda = self.cuda
        tr = self.tr
        if tr is not None:
            tr.begin_run()
        if cuda:
            cur = torch.cuda.current_stream(self.dev)
            self.s_update.wait_stream(cur)
            self.s_panel.wait_stream(cur)
        ev_diag = None     # tile (k,k) has all its updates      -> POTRF(k) may start
        ev_col = None      # column k has all its updates        -> TRSM(k) may start
        ns = self.nslots
        ev_upd = [None] * ns   # update k finished reading panel slot k % nslots
        # SM partition (tail of a factorization only): steps >= k_sw update on the `rest` SMs, and the POTRFs
        # that overlap those updates (steps > k_sw) run on the panel group
        part = cuda and factor and self.lookahead and self.s_potrf is not None
        k_sw = self._tail_start() if part else nt + 1
        self._su = self.s_update
        self._potrf_on_partition = False
        if part and k_sw < nt:
            self.s_rest.wait_stream(cur)
            self.s_potrf.wait_stream(cur)
        for k in range(nt):
            if part and k == k_sw:
                self.s_rest.wait_stream(self.s_update)
                self._su = self.s_rest
            self._potrf_on_partition = part and k > k_sw
            su = self._su
            # ---- panel k
            if cuda:
                with torch.cuda.stream(self.s_panel):
                    if ev_upd[k % ns] is not None:
                        self.s_panel.wait_event(ev_upd[k % ns])
                    if ev_diag is not None:
                        self.s_panel.wait_event(ev_diag)
                    if k == 0 and step0_gates:
                        self.s_panel.wait_event(step0_gates[0])      # columns 0 and 1 are on the device
                    self._panel_potrf(k, factor)
                    if ev_col is not None:
                        self.s_panel.wait_event(ev_col)
                    self._panel_rest(k, factor)
                    ev_panel = torch.cuda.Event()
                    ev_panel.record(self.s_panel)
                su.wait_event(ev_panel)
                if tr is not None and (k % self.grid.Q) == self.lay.q:
                    # this rank's TRSM k is enqueued: the L_kk slot it read may be overwritten
                    self.s_credit.wait_event(ev_panel)
                    tr.release_diag(k, self.s_credit.cuda_stream)
                if post_panel is not None:
                    post_panel(k, ev_panel)
            else:
                self._panel_potrf(k, factor)
                self._panel_rest(k, factor)
                if tr is not None and (k % self.grid.Q) == self.lay.q:
                    tr.release_diag(k, 0)
            # ---- trailing update k
            st = self._stream_ptr(su)
            if pre_update is not None:
                pre_update(k, st)
            off, nd, na, ntot = self.step_tasks[k]
            base = update_tasks_ptr + off * 32
            if not cuda:
                if ntot:
                    self._k_update(base, ntot, st)
                if tr is not None:
                    tr.release_panel(k, 0)
                continue
            with torch.cuda.stream(su):
                gated = k == 0 and bool(step0_gates)
                lazy = bool(step0_gates) and factor and self.lookahead and self.lazy_steps > 1
                if lazy and k == self.lazy_steps:
                    # the later upload groups catch up: updates 0 .. S-1, in order, group by group as they arrive
                    for g in range(1, len(self.step0_groups)):
                        su.wait_event(step0_gates[1 + g])
                        t0, t1, _, _ = self.step0_groups[g]
                        o0, _, na0, _ = self.step_tasks[0]
                        if t1 > t0:
                            self._k_update(update_tasks_ptr + (o0 + na0 + t0) * 32, t1 - t0, st)
                        for kk in range(1, self.lazy_steps):
                            lo_, cnt_ = self.lazy_index[(kk, g)]
                            if cnt_:
                                self._k_update(self.d_lazy.data_ptr() + lo_ * 32, cnt_, st)
                if gated:
                    for ev in (step0_gates if not self.lookahead else step0_gates[:1]):
                        su.wait_event(ev)
                if not self.lookahead:
                    if ntot:
                        self._k_update(base, ntot, st)
                    ev_diag = ev_col = ev_upd[k % ns] = self._record()
                    self._release_panel(k, ev_upd[k % ns])
                    continue
                # lookahead: (1) the diagonal tile of column k+1 so POTRF(k+1) can start, (2) the rest
                # of column k+1 so TRSM(k+1) can start, (3) everything else, overlapped with panel k+1.
                # A rank that owns nothing of a stage has nothing to wait for: it only receives, and may
                # join the broadcasts as soon as its receive slot is free (ev_upd of step k-1).
                ev_diag = ev_col = None
                self._mark("upd0", k, su)
                if nd:
                    self._k_update(base, nd, st)
                    ev_diag = self._record()
                if na > nd:
                    self._k_update(base + nd * 32, na - nd, st)
                if na:
                    ev_col = self._record()
                    self._mark("upd_a", k, su)
                if gated:
                    # host-resident input: release part b group by group behind the upload (lazy first steps:
                    # only the first group now, the others catch up before step S)
                    for g, (t0, t1, _, _) in enumerate(self.step0_groups[:1] if lazy else self.step0_groups):
                        su.wait_event(step0_gates[1 + g])
                        self._k_update(base + (na + t0) * 32, t1 - t0, st)
                elif lazy and k < self.lazy_steps:
                    lo_, cnt_ = self.lazy_index[(k, 0)]
                    if cnt_:
                        self._k_update(self.d_lazy.data_ptr() + lo_ * 32, cnt_, st)
                elif ntot > na:
                    # the bulk of the update overlaps panel step k+1: once it is small enough for that panel
                    # chain to bound the step, it runs at one CTA per SM and leaves the chain's kernels room
                    self._k_update(base + na * 32, ntot - na, st, thin=0 < ntot - na <= self.thin_tasks)
                ev_upd[k % ns] = self._record()
                self._mark("upd1", k, su)
                self._release_panel(k, ev_upd[k % ns])
        if cuda:
            cur.wait_stream(self.s_update)
            cur.wait_stream(self.s_panel)
            if part and k_sw < nt:
                cur.wait_stream(self.s_rest)
                cur.wait_stream(self.s_potrf)
            self._su = self.s_update
            self._potrf_on_partition = False
            if tr is not None:
                for s_ in self.s_sends.values():
                    cur.wait_stream(s_)
                cur.wait_stream(self.s_credit)
                tr.end_run(cur.cuda_stream)
        elif tr is not None:
            tr.end_run(0)

    def _mark(self, name: str, k: int, stream) -> None:
        """Timeline marker (only when self.trace is a list): a timing event on `stream`."""
        if self.trace is not None and self.cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            self.trace.append((name, k, ev))

    def trace_ms(self) -> list:
        """[(name, k, milliseconds since the first marker)] of the traced pass; synchronises."""
        torch.cuda.synchronize(self.dev)
        t0 = self.trace[0][2]
        return [(n, k, t0.elapsed_time(ev)) for n, k, ev in self.trace]

    def _release_panel(self, k: int, ev) -> None:
        """Update k is enqueued (event `ev`): tell the owners that panel slot k % nslots is free."""
        if self.tr is not None:
            self.s_credit.wait_event(ev)
            self.tr.release_panel(k, self.s_credit.cuda_stream)

    def _record(self):
        ev = torch.cuda.Event()
        ev.record(self._su or self.s_update)
        return ev

    def factor(self) -> None:
        """Enqueue the whole factorization (asynchronous on CUDA).  A <- L (lower tiles)."""
        self.d_info.zero_()
        self._run(self.d_tasks.data_ptr(), factor=True)

    def factor_from_host(self, host_in: torch.Tensor, host_out: torch.Tensor | None = None) -> None:
        """End-to-end form for callers whose tiles live in HOST memory (the tile-worker situation,
        worker_distrib.cpp:186,261): `host_in` is a pinned CPU tensor shaped like ``A.buf`` holding
        this rank's tiles; they are copied to the device, factored, and the factor is copied back
        into `host_out` (default: `host_in`).  Panel column k is final once its panel step is done,
        so its device-to-host copy runs on a copy stream underneath the remaining updates.
        Asynchronous; synchronise the current stream before reading `host_out`."""
        if not self.cuda:
            raise RuntimeError("factor_from_host needs a CUDA device")
        host_out = host_in if host_out is None else host_out
        assert host_in.shape == self.A.buf.shape and host_in.is_pinned() and host_out.is_pinned()
        cur = torch.cuda.current_stream(self.dev)
        if getattr(self, "s_copy", None) is None:
            self.s_copy = torch.cuda.Stream(self.dev)
        if getattr(self, "s_h2d", None) is None:
            self.s_h2d = torch.cuda.Stream(self.dev)
        self.d_info.zero_()
        # upload in storage (column-major) order on its own stream: columns 0-1 first, then the
        # groups of step 0's part b; each group of the update waits only for its own tiles
        gates = []
        self.s_h2d.wait_stream(cur)
        with torch.cuda.stream(self.s_h2d):
            for lo, hi in [(0, self.step0_head)] + [(g[2], g[3]) for g in self.step0_groups]:
                if hi > lo:
                    self.A.buf[lo:hi].copy_(host_in[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.s_h2d)
                gates.append(ev)
        lay, nt = self.lay, self.nt
        bounds = {j: (lay.col_start[j], lay.col_start[lay.cols[n + 1]] if n + 1 < len(lay.cols) else lay.ntiles)
                  for n, j in enumerate(lay.cols)}

        def post_panel(k: int, ev_panel) -> None:
            if k in bounds and bounds[k][1] > bounds[k][0]:
                lo, hi = bounds[k]
                self.s_copy.wait_event(ev_panel)
                with torch.cuda.stream(self.s_copy):
                    host_out[lo:hi].copy_(self.A.buf[lo:hi], non_blocking=True)

        self._run(self.d_tasks.data_ptr(), factor=True, post_panel=post_panel, step0_gates=gates)
        cur.wait_stream(self.s_copy)
        cur.wait_stream(self.s_h2d)

    def info(self) -> int:
        """LAPACK info of the last factor(): 0, or the 1-based global index of the first
        non-positive pivot (v6_test.c:56,95).  Synchronises; with several ranks the smallest
        non-zero value over the ranks is returned on every rank."""
        v = int(self.d_info.item())
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([v if v > 0 else 2 ** 31 - 1], dtype=torch.int64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
            v = int(t.item())
            v = 0 if v == 2 ** 31 - 1 else v
        return v

    # ---- residual: R = A0 - L L^T with the same schedule ---------------------------------------
    def residual(self, A0: TileMatrix) -> dict:
        """Overwrite A0 (same geometry as A, holding the original matrix) with the lower tiles of
        R = A0 - L L^T and return {"fro": ||R||_F/||A0||_F, "inf": ||R||_inf/||A0||_inf}: what
        v6_test.c:72-86 means to print (its dlauum forms L^T L instead).  Runs the
        factorization's own broadcast/update schedule with C pointing into A0."""
        assert A0.desc == self.A.desc and A0.rank == self.rank
        b, nt, tb, lay, g = self.b, self.nt, self.tile_bytes, self.lay, self.grid
        na_f, na_i = _norms(A0, self)
        # tril(L_kk) for every k this rank needs: broadcast from the diag owner to everyone
        tk = torch.empty((2, b, b), dtype=torch.float64, device=self.dev)
        shift = A0.buf.data_ptr() - self.A.buf.data_ptr()
        tasks = self.tasks_host.copy()
        tasks[:, 0] += shift
        d_tasks = torch.from_numpy(tasks).to(self.dev) if self.n_update_tasks else self.d_tasks
        col_tasks = {}
        for k in range(nt):
            if (k % g.Q) != lay.q:
                continue
            rows = np.asarray(lay.rows_in_col(k, k - 1), dtype=np.int64)
            if rows.size == 0:
                continue
            rec = np.empty((rows.size, 4), dtype=np.int64)
            rec[:, 0] = [A0.tile_ptr(int(i), k) for i in rows]
            rec[:, 1] = [tk[k % 2].data_ptr() if i == k else self.A.tile_ptr(int(i), k) for i in rows]
            rec[:, 2] = tk[k % 2].data_ptr()
            rec[:, 3] = (rows == k).astype(np.int64)
            col_tasks[k] = torch.from_numpy(rec).to(self.dev)

        def pre_update(k: int, st: int) -> None:
            # T_k = tril(L_kk) on its owner, sent down the process column, then column k of R
            kq, kp = k % g.Q, k % g.P
            if kq != lay.q:
                return
            ctx = torch.cuda.stream(self.s_update) if self.cuda else _Null()
            with ctx:
                if kp == lay.p:
                    self._k_tril(self.A.tile_ptr(k, k), tk[k % 2].data_ptr(), st)
                if g.P > 1:
                    self._bcast(tk[k % 2], kp, self._column_group(kq))
                if k in col_tasks:
                    self._k_update(col_tasks[k].data_ptr(), col_tasks[k].shape[0], st)

        self._run(d_tasks.data_ptr(), factor=False, pre_update=pre_update)
        nr_f, nr_i = _norms(A0, self)
        return {"fro": nr_f / na_f, "inf": nr_i / na_i, "norm_fro": na_f, "norm_inf": na_i}


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _norms(M: TileMatrix, ch: TiledCholesky) -> tuple[float, float]:
    """Frobenius and infinity norm of the symmetric matrix whose lower tiles are M (the strict
    upper triangle of diagonal tiles is ignored).  Plain torch reductions on the tile buffer:
    off the timed path (v6_test.c:72,84 dlange)."""
    b, nt = M.b, M.nt
    rows = torch.zeros(nt * b, dtype=torch.float64, device=M.device)
    ssq = torch.zeros((), dtype=torch.float64, device=M.device)
    for i, j in M.layout.tiles():
        t = M.tile(i, j)  # t[c, r] = A(r, c)
        if i == j:
            low = torch.triu(t)          # torch-upper of the transposed view == lower triangle of the tile
            strict = torch.triu(t, 1)
            ssq += (low * low).sum() + (strict * strict).sum()
            a = low.abs()
            rows[i * b:(i + 1) * b] += a.sum(0) + strict.abs().sum(1)
        else:
            ssq += 2.0 * (t * t).sum()
            a = t.abs()
            rows[i * b:(i + 1) * b] += a.sum(0)   # sum over columns c -> per row r of block i
            rows[j * b:(j + 1) * b] += a.sum(1)   # mirrored tile (j,i): per column c
    if ch.world > 1:
        import torch.distributed as dist
        dist.all_reduce(rows, group=ch.group)
        dist.all_reduce(ssq, group=ch.group)
    return float(ssq.sqrt().item()), float(rows.max().item())


def potrf_tile_desc(uplo: str, A: TileMatrix, group=None, lookahead: bool = True) -> int:
    """int CHAMELEON_dpotrf_Tile(cham_uplo_t uplo, CHAM_desc_t *A) (v6_test.c:56): in-place Cholesky of
    the tiled SPD matrix; returns LAPACK info.  uplo = 'L' (ChamLower, the one the reference calls): A = L L^T,
    storage position (i, j), i >= j, holds tile (i, j).  uplo = 'U' (ChamUpper): A = U^T U, and storage position
    (i, j), i >= j, holds the UPPER tile (j, i) (block row j, block column i) of A on entry and of U on exit;
    only the upper triangle of the diagonal tiles is referenced.  Since U = L^T, the upper case transposes
    every tile in place, runs the lower factorization and transposes back (two sweeps over the matrix at HBM
    speed next to N^3/3 flops)."""
    if uplo in ("U", "u"):
        if A.device.type != "cuda":
            raise RuntimeError("potrf_tile_desc('U') needs a CUDA device: there is no CPU path")
        transpose_tiles(A)
        try:
            return potrf_tile_desc("L", A, group=group, lookahead=lookahead)
        finally:
            transpose_tiles(A)
    if uplo not in ("L", "l"):
        raise ValueError("uplo must be 'L' (ChamLower) or 'U' (ChamUpper)")
    # one plan (and, on several ranks, one set of communicators / peer buffers) per descriptor storage:
    # repeated calls on the same matrix must not create new process groups every time
    key = (A.desc, A.rank, A.buf.data_ptr(), lookahead, id(group))
    ch = _plans.get(key)
    if ch is None:
        if len(_plans) >= 8:
            _, old = _plans.popitem()
            old.close()
        ch = _plans[key] = TiledCholesky(A, group=group, lookahead=lookahead)
    ch.factor()
    return ch.info()


def transpose_tiles(A: TileMatrix) -> None:
    """Transpose every local tile of A in place (asynchronous on the current stream)."""
    b = A.b
    st = torch.cuda.current_stream(A.device).cuda_stream
    _lib.call("chol_tile_transpose", b, A.buf.data_ptr(), b, b * b, A.buf.shape[0], st)


_plans: dict = {}
_partitions: dict = {}      # device index -> (chol_partition_t, panel stream, rest stream) | None, one per process
