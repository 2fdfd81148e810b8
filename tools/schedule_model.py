"""Critical-path model of the multi-GPU schedule (development tool, CPU only).

Builds the whole-job dependency graph with tests/_schedule_sim.py (the real planner and the real
`_run`, fake streams), puts the kernel durations measured on B200 (profiles/) on
the nodes and reports the longest path = predicted time of one factorization.  Used to see where an
8-GPU run loses time and what a change of the panel chain or of the transport would buy before
spending GPU minutes on it.

    python tools/schedule_model.py [--nt 64] [--bcast-gbs 150] [--update-tflops 33.3] ...
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _schedule_sim as S  # noqa: E402
from dense_linear_app_b200.tiles import TileDesc, TileMatrix  # noqa: E402


def model(P, Q, nt, a, transport="nccl"):
    b_sim = 16
    N = nt * b_sim
    mk = lambda r: TileMatrix(TileDesc(b_sim, b_sim, b_sim * b_sim, N, N, 0, 0, N, N, P, Q), r, "cpu")  # noqa: E731
    g, problems, _ = S.simulate(mk, P * Q, lookahead=True, transport=transport)
    assert not problems, problems
    tile_bytes = a.tile ** 2 * 8
    t_task = 2 * a.tile ** 3 / (a.update_tflops * 1e12)         # one GEMM tile update
    dur = [0.0] * len(g.kind)
    for n, kind in enumerate(g.kind):
        acc = g.acc[n]
        if kind == "potrf":
            dur[n] = a.potrf_ms * 1e-3
        elif kind == "trsm":
            ntiles = sum(1 for (reg, w) in acc if w)
            dur[n] = max(a.trsm_min_ms, a.trsm_tile_ms * ntiles + 0.12) * 1e-3
        elif kind == "update":
            ntasks = sum(1 for (reg, w) in acc if w)
            dur[n] = max(ntasks * t_task, 1.35e-4) + 1e-5
        elif kind == "bcast-done":
            ntiles = max(1, len({reg for reg, _ in acc}))
            dur[n] = a.bcast_lat_us * 1e-6 + ntiles * tile_bytes / (a.bcast_gbs * 1e9)
        elif kind == "peer-copy":
            ntiles = sum(1 for (reg, w) in acc if w)
            dur[n] = 5e-6 + ntiles * tile_bytes / (a.p2p_gbs * 1e9)
        elif kind in ("post", "flagwait"):
            dur[n] = 3e-6
    # longest path (nodes are created in a topological-compatible order per rank, but cross-rank edges
    # need a real topological pass)
    npred = [len(p) for p in g.preds]
    succ = [[] for _ in g.kind]
    for bnode, ps in enumerate(g.preds):
        for p_ in ps:
            succ[p_].append(bnode)
    ready = [i for i, c in enumerate(npred) if c == 0]
    finish = [0.0] * len(g.kind)
    start = [0.0] * len(g.kind)
    for i in ready:
        finish[i] = start[i] + dur[i]
        for s_ in succ[i]:
            start[s_] = max(start[s_], finish[i])
            npred[s_] -= 1
            if npred[s_] == 0:
                ready.append(s_)
    total = max(finish)
    busy = {}
    for n, kind in enumerate(g.kind):
        if kind == "update":
            busy[g.rank[n]] = busy.get(g.rank[n], 0.0) + dur[n]
    return total, max(busy.values()), sum(busy.values()) / len(busy)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nt", type=int, default=64)
    ap.add_argument("--tile", type=int, default=1024)
    ap.add_argument("--update-tflops", type=float, default=33.3, help="update kernel per GPU (1-GPU measurement)")
    ap.add_argument("--potrf-ms", type=float, default=0.96)
    ap.add_argument("--trsm-min-ms", type=float, default=0.45)
    ap.add_argument("--trsm-tile-ms", type=float, default=0.048)
    ap.add_argument("--bcast-gbs", type=float, default=150.0)
    ap.add_argument("--bcast-lat-us", type=float, default=40.0)
    ap.add_argument("--p2p-gbs", type=float, default=300.0)
    a = ap.parse_args()
    flops = (a.nt * a.tile) ** 3 / 3
    print(f"N={a.nt * a.tile} tile={a.tile}: critical-path model (update {a.update_tflops} TFLOP/s per GPU)")
    t1 = None
    for (P, Q) in ((1, 1), (1, 2), (2, 2), (2, 4)):
        for tr in (("nccl",) if P * Q == 1 else ("nccl", "peer")):
            t, busy_max, busy_avg = model(P, Q, a.nt, a, tr)
            t1 = t1 or t
            print(f"  {P}x{Q} {tr:5s}: {t * 1e3:8.1f} ms  {flops / t / 1e12:7.1f} TFLOP/s  efficiency {t1 / (P * Q) / t * 100:5.1f} %"
                  f"   update busy max/avg per rank {busy_max * 1e3:7.1f}/{busy_avg * 1e3:7.1f} ms")


if __name__ == "__main__":
    main()
