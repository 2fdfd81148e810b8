"""Multi-GPU parity on hardware: every rank's tiles of the P x Q block-cyclic
factorization against the SAME matrix factored on one GPU (1 x 1 grid, computed by every rank on
its own GPU), for the peer-push transport and the NCCL transport, plus a timing of both.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/mgpu_parity.py [--sizes 4096:512,16384:1024] [--time 32768:1024]

Rank 0 prints one JSON line per case (also appended to gpurun_out/mgpu_parity_<N>gpu.jsonl).
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dense_linear_app_b200 import runtime  # noqa: E402
from dense_linear_app_b200.cholesky import TiledCholesky  # noqa: E402
from dense_linear_app_b200.grid import ProcessGrid  # noqa: E402
from dense_linear_app_b200.tiles import TileDesc, TileMatrix  # noqa: E402


def factor_grid(N, b, g, rank, transport, reps=1, lookahead=True):
    import torch.distributed as dist
    os.environ["CHOL_PANEL_TRANSPORT"] = transport
    desc = TileDesc(b, b, b * b, N, N, 0, 0, N, N, g.P, g.Q)
    M = TileMatrix(desc, rank).generate(float(N), 42)
    pristine = M.buf.clone()
    ch = TiledCholesky(M, lookahead=lookahead)
    times = []
    for rep in range(reps + 1):
        M.buf.copy_(pristine)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ch.factor()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=M.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rep:
            times.append(float(t.item()))
    info = ch.info()
    A0 = TileMatrix(desc, rank)
    A0.buf.copy_(pristine)
    res = ch.residual(A0)
    recv = ch.tr.bytes_received_per_run() if ch.tr is not None else None
    ch.close()
    del A0, pristine
    return M, info, res, times, recv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="2048:256,4096:512,16384:1024")
    ap.add_argument("--time", default="32768:1024")
    ap.add_argument("--transports", default="peer,nccl")
    ap.add_argument("--time-transports", default=None)
    a = ap.parse_args()
    import torch.distributed as dist
    rank, world = runtime.init()
    g = ProcessGrid.for_world(world)
    if os.environ.get("CHOL_GRID"):
        gp, gq = (int(x) for x in os.environ["CHOL_GRID"].lower().split("x"))
        g = ProcessGrid(gp, gq)
    dev = torch.device("cuda", torch.cuda.current_device())
    out_path = os.path.join(ROOT, "gpurun_out", f"mgpu_parity_{world}gpu.jsonl")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)

    def emit(rec):
        if rank == 0:
            line = json.dumps(rec)
            print(line, flush=True)
            with open(out_path, "a") as f:
                f.write(line + "\n")

    transports = a.transports.split(",")
    for case in [c for c in a.sizes.split(",") if c]:
        N, b = (int(x) for x in case.split(":"))
        # the same matrix on ONE GPU (every rank computes it on its own device)
        M1 = TileMatrix(TileDesc.square(N, b), 0, dev).generate(float(N), 42)
        ch1 = TiledCholesky(M1)
        ch1.factor()
        assert ch1.info() == 0
        lmax = float(M1.buf.abs().max().item())
        results = {}
        for tr in transports:
            try:
                M, info, res, _, recv = factor_grid(N, b, g, rank, tr, reps=0)
                worst, exact = 0.0, True
                for i, j in M.layout.tiles():
                    mine, ref = M.tile(i, j), M1.tile(i, j)
                    if i == j:            # strict upper of diagonal tiles is not part of L
                        mine, ref = torch.triu(mine), torch.triu(ref)   # torch view is the transpose
                    d = float((mine - ref).abs().max().item())
                    worst = max(worst, d)
                    exact = exact and bool(torch.equal(mine, ref))
                t = torch.tensor([worst, 0.0 if exact else 1.0], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                results[tr] = {"info": info, "max_abs_diff_vs_1gpu_over_max_L": float(t[0].item()) / lmax,
                               "bit_identical_to_1gpu": bool(t[1].item() == 0.0), "backward_error": res["fro"],
                               "residual_inf": res["inf"], "bytes_received_rank0": recv}
                results[tr]["ok"] = info == 0 and results[tr]["max_abs_diff_vs_1gpu_over_max_L"] <= 1e-13 \
                    and res["fro"] <= 1e-13
                del M
            except Exception as e:  # noqa: BLE001
                results[tr] = {"error": repr(e)}
                raise
        emit({"case": "parity", "N": N, "tile": b, "grid": f"{g.P}x{g.Q}", "n_gpus": world, "results": results})
        del M1, ch1
        torch.cuda.empty_cache()
    for case in [c for c in a.time.split(",") if c]:
        N, b = (int(x) for x in case.split(":"))
        rec = {"case": "timing", "N": N, "tile": b, "grid": f"{g.P}x{g.Q}", "n_gpus": world}
        for tr in (a.time_transports.split(",") if a.time_transports else transports):
            t0 = time.time()
            M, info, res, times, recv = factor_grid(N, b, g, rank, tr, reps=3)
            best = min(times)
            rec[tr] = {"ms_best": best, "ms_all": times, "tflops_best": N ** 3 / 3 / (best * 1e-3) / 1e12,
                       "info": info, "backward_error": res["fro"], "wall_s": time.time() - t0,
                       "bytes_received_rank0": recv}
            del M
            torch.cuda.empty_cache()
        emit(rec)
    runtime.finalize()


if __name__ == "__main__":
    main()
