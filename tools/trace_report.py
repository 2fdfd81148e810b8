"""Summarise traces written by tools/mgpu_trace.py: per step the POTRF / TRSM durations, the gaps on
the critical chain and the update stages.  usage: python tools/trace_report.py gpurun_out/trace_8gpu_N65536_rank*.json"""
import json
import sys

ranks = {}
for path in sys.argv[1:]:
    r = int(path.split("rank")[1].split(".")[0])
    ev = {}
    for name, k, ms in json.load(open(path)):
        ev[(name, k)] = ms
    ranks[r] = ev
nt = 1 + max(k for ev in ranks.values() for (_, k) in ev)
end = max(ms for ev in ranks.values() for ms in ev.values())
print(f"ranks {sorted(ranks)}  steps {nt}  last marker at {end:.2f} ms")
print(" k | potrf ms | trsm ms (max) | potrf0->potrf0(k+1) | upd stage-a max | upd total max | end of upd(k) max")
prev = None
tot = {"potrf": 0, "trsm": 0}
for k in range(nt):
    p0 = [ev[("potrf0", k)] for ev in ranks.values() if ("potrf0", k) in ev]
    p1 = [ev[("potrf1", k)] for ev in ranks.values() if ("potrf1", k) in ev]
    t = [ev[("trsm1", k)] - ev[("trsm0", k)] for ev in ranks.values() if ("trsm0", k) in ev]
    ua = [ev[("upd_a", k)] - ev[("upd0", k)] for ev in ranks.values() if ("upd_a", k) in ev]
    ut = [ev[("upd1", k)] - ev[("upd0", k)] for ev in ranks.values() if ("upd1", k) in ev]
    ue = [ev[("upd1", k)] for ev in ranks.values() if ("upd1", k) in ev]
    chain = (p0[0] - prev) if (prev is not None and p0) else float("nan")
    if p0:
        prev = p0[0]
    tot["potrf"] += (p1[0] - p0[0]) if p0 else 0
    tot["trsm"] += max(t) if t else 0
    if k < 4 or k % 4 == 0 or k > nt - 12:
        print(f"{k:3d} | {(p1[0] - p0[0]) if p0 else float('nan'):8.3f} | {max(t) if t else float('nan'):8.3f} | {chain:8.3f} | "
              f"{max(ua) if ua else float('nan'):8.3f} | {max(ut) if ut else float('nan'):8.3f} | {max(ue) if ue else float('nan'):9.2f}")
print("sum potrf %.2f ms, sum max trsm %.2f ms" % (tot["potrf"], tot["trsm"]))
