"""TEST INFRASTRUCTURE: whole-job simulation of the multi-rank schedule on the CPU.

Every rank's real ``TiledCholesky._run`` is executed with fake CUDA streams/events; kernels,
collectives, peer copies and flags become nodes of ONE dependency graph over all ranks:

  * stream order, event record -> wait, stream -> stream waits;
  * an NCCL-style broadcast: every participant posts (node) and completes (node); a completion
    depends on every participant's post.  Collectives are matched per communicator by call order,
    a missing or mis-ordered participant is reported as a hang;
  * the peer-push transport (transport.py) with its primitives recorded: a push is a node on the send
    stream that reads the owner's tiles and WRITES the reader's slot; a flag wait depends on the post
    of the same (rank, word, value) — a wait whose post never happens is reported as a hang.

Checks: the graph is acyclic (no deadlock), and any two accesses to the same tile/buffer of the same
rank with at least one write are ordered by reachability (no race) — including a peer's remote write
into a receive slot against the local update still reading it.
"""
import contextlib
import ctypes as C

import numpy as np
import torch

from dense_linear_app_b200.cholesky import TiledCholesky
from dense_linear_app_b200.transport import PeerTransport


class Graph:
    def __init__(self):
        self.kind, self.rank, self.preds, self.acc = [], [], [], []   # acc[n] = [(region, is_write)]
        self.coll = {}      # (comm key, seq) -> [(rank, post, done)]
        self.comm_size = {}
        self.posts, self.waits = {}, {}   # (rank, word, value) -> node / [nodes]

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
        for key, ws in self.waits.items():
            if key not in self.posts:
                if key[2] != 0:                 # value 0 = the zero-filled initial state
                    problems.append(f"flag wait {key} (rank, word, value) is never posted (hang)")
                continue
            for w in ws:
                self.edge(self.posts[key], w)
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


STREAMS = {}


class SimStream:
    def __init__(self, g, rank):
        self.g, self.rank, self.last = g, rank, None
        self.cuda_stream = id(self)
        STREAMS[self.cuda_stream] = self

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


class SimTransport(PeerTransport):
    """transport.py with its primitives turned into graph nodes.  All ranks live in one process, so
    the "peer-mapped" addresses are simply the addresses of the other ranks' buffers."""

    def __init__(self, sim, *a, nslots=None, credits=True, **k):
        self.sim, self.credits = sim, credits
        if nslots:
            self.NSLOTS = nslots
        super().__init__(*a, **k)

    def _open(self):
        self.buf = torch.zeros(self.nbytes // 8 + 1, dtype=torch.float64)
        self.local = self.buf.data_ptr()

    def connect(self, sims):
        self.peer_base = {s.rank: s.tr.local for s in sims if s.rank != self.rank}

    def _send(self, pushes, ready_stream, send_stream_of):
        g, sim = self.sim.g, self.sim
        ready = STREAMS[ready_stream].last
        for p in pushes:
            st = STREAMS[send_stream_of(p.peer)]
            rd = {sim.region(p.src + t * p.src_stride * p.tile_bytes) for t in range(p.count)}
            wr = {sim.region(self.peer_base[p.peer] + p.dst_off + t * p.dst_stride * p.tile_bytes)
                  for t in range(p.count)}
            n = st.chain(g.node("peer-copy", self.rank, rd, wr))
            g.edge(ready, n)
            if p.credit is not None and self.credits:
                g.waits.setdefault((self.rank, p.credit, p.credit_value), []).append(n)
            if p.flag is not None:
                g.posts[(p.peer, p.flag, p.flag_value)] = st.chain(g.node("post", self.rank))

    def _wait(self, flag, value, stream):
        st = STREAMS[stream]
        self.sim.g.waits.setdefault((self.rank, flag, value), []).append(st.chain(self.sim.g.node("flagwait", self.rank)))

    def _post(self, targets, value, stream):
        st = STREAMS[stream]
        n = st.chain(self.sim.g.node("post", self.rank))
        for r, f in targets:
            self.sim.g.posts[(r, f, value)] = n


class RankSim(TiledCholesky):
    """One rank of the simulated job.  `world_sims` (filled by simulate()) gives access to the peers'
    receive buffers for the symmetric transport."""

    def __init__(self, g, A, lookahead=True, transport="nccl", nslots=None, credits=True):
        self.g = g
        self.A, self.nt, self.b = A, A.nt, A.b
        self.grid, self.rank, self.lay = A.grid, A.rank, A.layout
        self.dev, self.world, self.cuda = A.device, A.grid.size, True
        self.group, self.lookahead, self.transport = None, lookahead, transport
        self.update_events = None
        self.tile_bytes = self.b * self.b * 8
        self.work = torch.zeros(16, dtype=torch.float64)
        self.d_info = torch.zeros(1, dtype=torch.int32)
        self.nslots, self.tr = 2, None
        if transport == "peer":
            self.tr = SimTransport(self, self.nt, self.b, self.grid, self.rank, self.work.numel(), nslots=nslots,
                                   credits=credits)
            self.nslots = self.tr.nslots
        self.panel = torch.zeros((2, max(self.nt - 1, 1), self.b, self.b), dtype=torch.float64)
        self.diag = torch.zeros((self.b, self.b), dtype=torch.float64)
        self._col_groups = [("col", q) for q in range(self.grid.Q)]
        g.comm_size["world"] = self.world
        for q in range(self.grid.Q):
            g.comm_size[("col", q)] = self.grid.P
        self.s_update, self.s_panel = SimStream(g, self.rank), SimStream(g, self.rank)
        self.s_sends = {r: SimStream(g, self.rank) for r in range(self.world) if r != self.rank}
        self.s_credit = SimStream(g, self.rank)
        self.cur = SimStream(g, self.rank)
        self._seq = {}
        self.world_sims = None

    # ---- addresses -> (rank, buffer, tile) regions
    def region(self, ptr):
        tb = self.tile_bytes
        for sim in (self.world_sims or [self]):
            for name, t in (("A", sim.A.buf), ("P", sim.panel), ("D", sim.diag), ("W", sim.work)):
                if t.data_ptr() <= ptr < t.data_ptr() + t.numel() * 8:
                    return (sim.rank, name, (ptr - t.data_ptr()) // tb if name in ("A", "P") else 0)
            tr = sim.tr
            if tr is not None and tr.local <= ptr < tr.local + tr.nbytes:
                off = ptr - tr.local
                if off >= tr.off_panel:
                    return (sim.rank, "P", (off - tr.off_panel) // tb)
                assert off >= tr.off_diag, "pointer into the flag words"
                return (sim.rank, "D", (off - tr.off_diag) // tr.diag_stride)
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

    def _k_trsm_panel(self, l_ptr, work_ptr, tiles_ptr, ntiles, st):
        tl = {self.region(p) for p in (C.c_int64 * ntiles).from_address(tiles_ptr)}
        self.op("trsm", tl | {self.region(l_ptr), self.region(work_ptr)}, tl)

    def _k_update(self, tasks_ptr, ntasks, st, thin=False):
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


def simulate(make_matrix, world, runs=1, **kw):
    """Run every rank's factor() schedule (`runs` times back to back) into one graph.  Returns
    (graph, problems, races)."""
    g = Graph()
    sims = [RankSim(g, make_matrix(r), **kw) for r in range(world)]
    for s in sims:
        s.world_sims = sims
    for s in sims:
        if s.tr is not None:
            s.tr.connect(sims)
        s._build_plan()            # needs every rank's buffers for the region lookup

    saved = (torch.cuda.Event, torch.cuda.Stream, torch.cuda.stream, torch.cuda.current_stream)
    torch.cuda.Event, torch.cuda.stream = SimEvent, stream_ctx
    torch.cuda.current_stream = lambda *a, **k: CUR[-1]
    try:
        for s in sims:
            CUR.clear()
            CUR.append(s.cur)
            for _ in range(runs):
                s._run(s.d_tasks.data_ptr(), factor=True)
    finally:
        torch.cuda.Event, torch.cuda.Stream, torch.cuda.stream, torch.cuda.current_stream = saved
        CUR.clear()
    problems = g.link()
    anc, cyc = g.ancestors()
    if cyc:
        return g, problems + [cyc], None
    return g, problems, g.races(anc)
