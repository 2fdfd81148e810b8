"""TEST INFRASTRUCTURE: whole-job simulation of the multi-rank schedule on the CPU.

Every rank's real ``TiledCholesky._run`` is executed with fake CUDA streams/events; kernels,
collectives, peer copies and flags become nodes of ONE dependency graph over all ranks:

  * stream order, event record -> wait, stream -> stream waits;
  * an NCCL-style broadcast: every participant posts (node) and completes (node); a completion
    depends on every participant's post.  Collectives are matched per communicator by call order,
    a missing or mis-ordered participant is reported as a hang;
  * symmetric-memory flags: the k-th ``put_signal(src -> dst)`` releases the k-th
    ``wait_signal(dst <- src)``; a put waits for the previous flag of the same pair to be consumed.

Checks: the graph is acyclic (no deadlock), and any two accesses to the same tile/buffer of the same
rank with at least one write are ordered by reachability (no race) — including a peer's remote write
into a receive slot against the local update still reading it.
"""
import contextlib
import ctypes as C

import numpy as np
import torch

from dense_linear_app_b200.cholesky import TiledCholesky


class Graph:
    def __init__(self):
        self.kind, self.rank, self.preds, self.acc = [], [], [], []   # acc[n] = [(region, is_write)]
        self.coll = {}      # (comm key, seq) -> [(rank, post, done)]
        self.comm_size = {}
        self.puts, self.waits = {}, {}   # (src, dst, ch) -> [node]

    def node(self, kind, rank, reads=(), writes=()):
        self.kind.append(kind)
        self.rank.append(rank)
        self.preds.append(set())
        self.acc.append([(r, False) for r in reads] + [(w, True) for w in writes])
        return len(self.kind) - 1

    def edge(self, a, b):
        if a is not None:
            self.preds[b].add(a)

    # ---- after every rank has been simulated
    def link(self):
        problems = []
        for (key, seq), parts in self.coll.items():
            if len(parts) != self.comm_size[key]:
                problems.append(f"collective #{seq} on {key}: {len(parts)} of {self.comm_size[key]} ranks take part (hang)")
            for _, post, _ in parts:
                for _, _, done in parts:
                    self.edge(post, done)
        for key in set(self.puts) | set(self.waits):
            p, w = self.puts.get(key, []), self.waits.get(key, [])
            if len(p) != len(w):
                problems.append(f"flags {key}: {len(p)} puts but {len(w)} waits")
            for k in range(min(len(p), len(w))):
                self.edge(p[k], w[k])
                if k + 1 < len(p):
                    self.edge(w[k], p[k + 1])
        return problems

    def ancestors(self):
        n = len(self.kind)
        indeg = [0] * n
        succ = [[] for _ in range(n)]
        for b, ps in enumerate(self.preds):
            for a in ps:
                succ[a].append(b)
                indeg[b] += 1
        order = [i for i in range(n) if indeg[i] == 0]
        anc = [0] * n
        for i in order:                     # list grows while iterating: Kahn's algorithm
            for b in succ[i]:
                anc[b] |= anc[i] | (1 << i)
                indeg[b] -= 1
                if indeg[b] == 0:
                    order.append(b)
        if len(order) != n:
            stuck = [f"{self.kind[i]}@r{self.rank[i]}" for i in range(n) if indeg[i] > 0][:8]
            return None, f"dependency cycle = deadlock, e.g. {stuck}"
        return anc, None

    def races(self, anc):
        by_region = {}
        for n, accs in enumerate(self.acc):
            for reg, w in accs:
                by_region.setdefault(reg, []).append((n, w))
        bad = []
        for reg, lst in by_region.items():
            for x in range(len(lst)):
                a, wa = lst[x]
                for y in range(x + 1, len(lst)):
                    b, wb = lst[y]
                    if a == b or not (wa or wb):
                        continue
                    if not ((anc[b] >> a) & 1 or (anc[a] >> b) & 1):
                        bad.append((reg, f"{self.kind[a]}@r{self.rank[a]}", f"{self.kind[b]}@r{self.rank[b]}"))
        return bad


class SimStream:
    def __init__(self, g, rank):
        self.g, self.rank, self.last = g, rank, None
        self.cuda_stream = id(self)

    def chain(self, n):
        self.g.edge(self.last, n)
        self.last = n
        return n

    def wait_event(self, ev):
        assert ev.node is not None, "waiting on an event that was never recorded"
        self.g.edge(ev.node, self.chain(self.g.node("wait", self.rank)))

    def wait_stream(self, other):
        n = self.chain(self.g.node("wait", self.rank))
        self.g.edge(other.last, n)


class SimEvent:
    def __init__(self, *a, **k):
        self.node = None

    def record(self, stream=None):
        s = stream or CUR[-1]
        self.node = s.chain(s.g.node("record", s.rank))


CUR = []


@contextlib.contextmanager
def stream_ctx(s):
    CUR.append(s)
    try:
        yield
    finally:
        CUR.pop()


class SymmHandle:
    """Fake of the _SymmetricMemory handle: flags only."""

    def __init__(self, g, rank):
        self.g, self.rank = g, rank

    def put_signal(self, dst, ch=0, timeout_ms=0):
        s = CUR[-1]
        self.g.puts.setdefault((self.rank, dst, ch), []).append(s.chain(self.g.node("put", self.rank)))

    def wait_signal(self, src, ch=0, timeout_ms=0):
        s = CUR[-1]
        self.g.waits.setdefault((src, self.rank, ch), []).append(s.chain(self.g.node("flagwait", self.rank)))


class RankSim(TiledCholesky):
    """One rank of the simulated job.  `world_sims` (filled by simulate()) gives access to the peers'
    receive buffers for the symmetric transport."""

    def __init__(self, g, A, lookahead=True, transport="nccl", nslots=None):
        self.g = g
        self.A, self.nt, self.b = A, A.nt, A.b
        self.grid, self.rank, self.lay = A.grid, A.rank, A.layout
        self.dev, self.world, self.cuda = A.device, A.grid.size, True
        self.group, self.lookahead, self.transport = None, lookahead, transport
        self.update_events = None
        self.tile_bytes = self.b * self.b * 8
        self.nslots = 2 if transport == "nccl" else (nslots or max(2, min(max(self.nt - 1, 1),
                                                                          self.grid.Q + self.grid.P + 2)))
        self.work = torch.zeros(16, dtype=torch.float64)
        self.d_info = torch.zeros(1, dtype=torch.int32)
        self.panel = torch.zeros((self.nslots, max(self.nt - 1, 1), self.b, self.b), dtype=torch.float64)
        self.diag = torch.zeros((self.b, self.b), dtype=torch.float64)
        self._col_groups = [("col", q) for q in range(self.grid.Q)]
        g.comm_size["world"] = self.world
        for q in range(self.grid.Q):
            g.comm_size[("col", q)] = self.grid.P
        self._build_plan()
        self.s_update, self.s_panel = SimStream(g, self.rank), SimStream(g, self.rank)
        self.s_sends = [SimStream(g, self.rank) if r != self.rank else None for r in range(self.world)]
        self.cur = SimStream(g, self.rank)
        self._symm = SymmHandle(g, self.rank)
        self._seq = {}
        self.world_sims = None

    @property
    def _peer_panel(self):
        return [s.panel if s.rank != self.rank else None for s in self.world_sims]

    # ---- addresses -> (rank, buffer, tile) regions
    def region(self, ptr):
        tb = self.tile_bytes
        for sim in (self.world_sims or [self]):
            for name, t in (("A", sim.A.buf), ("P", sim.panel), ("D", sim.diag), ("W", sim.work)):
                if t.data_ptr() <= ptr < t.data_ptr() + t.numel() * 8:
                    return (sim.rank, name, (ptr - t.data_ptr()) // tb if name in ("A", "P") else 0)
        raise AssertionError("pointer outside every known buffer")

    def regions_of(self, t):
        r, name, first = self.region(t.data_ptr())
        if name in ("D", "W"):
            return {(r, name, 0)}
        n = max(1, (t.numel() * 8 + self.tile_bytes - 1) // self.tile_bytes)
        return {(r, name, first + i) for i in range(n)}

    def op(self, kind, reads, writes):
        s = CUR[-1]
        return s.chain(self.g.node(kind, self.rank, reads, writes))

    # ---- recorded entry points
    def _potrf_workspace(self, b):
        return 8

    def _k_potrf(self, a_ptr, info_base, st):
        t = self.region(a_ptr)
        self.op("potrf", {t}, {t, (self.rank, "W", 0)})

    def _k_trsm_panel(self, l_ptr, tiles_ptr, ntiles, st):
        tl = {self.region(p) for p in (C.c_int64 * ntiles).from_address(tiles_ptr)}
        self.op("trsm", tl | {self.region(l_ptr), (self.rank, "W", 0)}, tl)

    def _k_update(self, tasks_ptr, ntasks, st):
        rec = np.ctypeslib.as_array((C.c_int64 * (4 * ntasks)).from_address(tasks_ptr)).reshape(ntasks, 4)
        w = {self.region(c) for c in rec[:, 0].tolist()}
        r = {self.region(p) for p in rec[:, 1].tolist()} | {self.region(p) for p in rec[:, 2].tolist()}
        self.op("update", r | w, w)

    def _bcast(self, t, src, group):
        key = group if isinstance(group, tuple) else "world"
        mine = (src == self.lay.p) if isinstance(group, tuple) else (src == self.rank)
        regs = self.regions_of(t)
        seq = self._seq.get(key, 0)
        self._seq[key] = seq + 1
        rd, wr = (regs, set()) if mine else (set(), regs)
        post = self.op("bcast-post", rd, wr)
        done = self.op("bcast-done", rd, wr)
        self.g.coll.setdefault((key, seq), []).append((self.rank, post, done))


def simulate(make_matrix, world, **kw):
    """Run every rank's factor() schedule into one graph.  Returns (graph, problems, races)."""
    g = Graph()
    sims = [RankSim(g, make_matrix(r), **kw) for r in range(world)]
    real_copy = torch.Tensor.copy_
    bufs = [(s.rank, s.panel.data_ptr(), s.panel.data_ptr() + s.panel.numel() * 8) for s in sims]
    for s in sims:
        s.world_sims = sims

    def recording_copy(dst, src, non_blocking=False):
        for r, lo, hi in bufs:
            if lo <= dst.data_ptr() < hi:                      # a peer copy into rank r's receive buffer
                sim = sims[CUR[-1].rank]
                sim.op("peer-copy", sim.regions_of(src), sim.regions_of(dst))
                return dst
        return real_copy(dst, src, non_blocking)

    saved = (torch.cuda.Event, torch.cuda.Stream, torch.cuda.stream, torch.cuda.current_stream, torch.Tensor.copy_)
    torch.cuda.Event, torch.cuda.stream = SimEvent, stream_ctx
    torch.cuda.current_stream = lambda *a, **k: CUR[-1]
    torch.Tensor.copy_ = recording_copy
    try:
        for s in sims:
            CUR.clear()
            CUR.append(s.cur)
            s._run(s.d_tasks.data_ptr(), factor=True)
    finally:
        torch.cuda.Event, torch.cuda.Stream, torch.cuda.stream, torch.cuda.current_stream, torch.Tensor.copy_ = saved
        CUR.clear()
    problems = g.link()
    anc, cyc = g.ancestors()
    if cyc:
        return g, problems + [cyc], None
    return g, problems, g.races(anc)
