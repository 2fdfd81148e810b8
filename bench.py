#!/usr/bin/env python
"""Headline benchmark: FP64 tiled Cholesky TFLOP/s (N^3/3) on 1..8 B200s, with backward error.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--N 65536] [--tile 1024]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One step = one factorization of the synthetic SPD matrix of BASELINE.json's configs[2]
(N=65536, tile 1024, dplgsy-style generator with bump=N, seed 42), resident in HBM; between steps
the factored tiles are overwritten from a pristine device copy (inside the timed region).  Rank 0
prints ONE JSON line.  `--impl reference` times the reference's own CPU Cholesky program
(oracle/_ref/lapacke_dpotrf_ref, compiled from the reference's lapacke_dpotrf.c) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fp64_cholesky_tflops"
UNIT = "TFLOP/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--N", type=int, default=65536)
    ap.add_argument("--tile", type=int, default=1024)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lookahead", action="store_true")
    return ap.parse_args()


def workload_config(a, P, Q):
    return {"workload": f"FP64 random SPD N={a.N}, tile {a.tile}, lower Cholesky A=LL^T (BASELINE configs[2])",
            "N": a.N, "tile": a.tile, "generator": "dplgsy-style LCG, bump=N, seed=42",
            "grid": f"{P}x{Q} block-cyclic", "flops": "N^3/3",
            "l2": "inputs larger than L2 (lower tiles %.1f GiB >> 126 MB)" % (a.N * (a.N + a.tile) / 2 * 8 / 2 ** 30)}


# ---- clocks ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v == "Active":
                    reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU reference arm ----------------------------------------------------------------------------------
def host_cores() -> int:
    return len(os.sched_getaffinity(0))


def run_cpu_reference_once(threads: int) -> dict:
    """One run of the reference's own CPU Cholesky (lapacke_dpotrf.c, N=12000 fixed in its source,
    timed around LAPACKE_dpotrf only, GF/s = N^3/3/t) — or, if oracle/_ref is missing, the oracle's
    OpenBLAS port on the same size."""
    from oracle import oracle as O
    if os.path.exists(O.REF_BIN):
        r = O.run_reference_binary(threads=threads)
        return {"kind": "reference", "N": r["N"], "seconds": r["seconds"], "tflops": r["gflops"] / 1e3,
                "sample": f"reference lapacke_dpotrf.c (oracle/_ref), N={r['N']}, LAPACKE_dpotrf on OpenBLAS, "
                          f"{threads} threads, one factorization"}
    N = 12000
    A = O.plgsy(float(N), N, 42)
    t0 = time.time()
    _, info = O.blas_potrf(A, threads=threads)
    dt = time.time() - t0
    assert info == 0
    return {"kind": "port", "N": N, "seconds": dt, "tflops": N ** 3 / 3 / dt / 1e12,
            "sample": f"oracle port: OpenBLAS dpotrf N={N}, {threads} threads, one factorization"}


def run_cpu_tiled_single_worker() -> dict:
    """BASELINE configs[0]: N=4096, tile 512, the reference-style tiled Cholesky on ONE CPU worker
    (tile DAG in the client's order, OpenBLAS tile kernels with 1 BLAS thread as benchmark.c:173-175
    forces) — the oracle's port of that path, timed; a reported baseline only."""
    from oracle import oracle as O
    N, b = 4096, 512
    A = O.plgsy(float(N), N, 42)
    tiles = O.to_tiles(A, b)
    t0 = time.time()
    info = O.blas_potrf_tiled(tiles, N // b, b, threads=1)
    dt = time.time() - t0
    L = O.from_tiles(tiles, N // b, b)
    return {"workload": "N=4096, tile 512, tile DAG, 1 worker, 1 BLAS thread (configs[0])", "seconds": dt,
            "tflops": N ** 3 / 3 / dt / 1e12, "info": int(info), "backward_error": O.backward_error_blas(A, L)}


def main_reference(a) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from dense_linear_app_b200.grid import ProcessGrid
    g = ProcessGrid.for_world(a.gpus)
    cores = host_cores()
    for _ in range(a.warmup):
        run_cpu_reference_once(cores)
    runs = [run_cpu_reference_once(cores) for _ in range(a.steps)]
    tf = sum(r["N"] ** 3 / 3 for r in runs) / sum(r["seconds"] for r in runs) / 1e12
    ms = sum(r["seconds"] for r in runs) / len(runs) * 1e3
    tiled = run_cpu_tiled_single_worker()
    line = {"impl": "reference", "metric": METRIC, "value": tf, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(a, g.P, g.Q),
            "cpu_baseline": {"value": tf, "unit": UNIT, "cores": cores, "kind": runs[0]["kind"],
                             "sample": runs[0]["sample"] + f" per step, {a.steps} steps",
                             "tiled_single_worker": tiled},
            "e2e": {"value": tf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---- our arm --------------------------------------------------------------------------------------------
def main_ours(a) -> int:
    import torch
    from dense_linear_app_b200 import _lib, runtime
    from dense_linear_app_b200.cholesky import TiledCholesky
    from dense_linear_app_b200.grid import ProcessGrid
    from dense_linear_app_b200.tiles import TileDesc, TileMatrix

    rank, world = runtime.init()
    if world != a.gpus:
        if rank == 0:
            sys.stderr.write(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run\n")
        return 2
    dist = torch.distributed if world > 1 else None
    g = ProcessGrid.for_world(world)
    if os.environ.get("CHOL_GRID"):            # e.g. CHOL_GRID=2x1: experiment with another grid shape
        gp, gq = (int(x) for x in os.environ["CHOL_GRID"].lower().split("x"))
        assert gp * gq == world
        g = ProcessGrid(gp, gq)
    dev = torch.device("cuda", torch.cuda.current_device())
    N, b = a.N, a.tile
    desc = TileDesc(b, b, b * b, N, N, 0, 0, N, N, g.P, g.Q)
    M = TileMatrix(desc, rank).generate(float(N), 42)
    pristine = M.buf.clone()
    ch = TiledCholesky(M, lookahead=not a.no_lookahead)
    lib = _lib.load()
    flops = float(N) ** 3 / 3.0

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        M.buf.copy_(pristine)
        ch.factor()

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    ch.update_events = []
    n0 = lib.chol_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    launches = lib.chol_launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    elapsed = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    elapsed = float(elapsed.item())
    upd = ch.update_events
    ch.update_events = None
    upd_s = sum(x.elapsed_time(y) for x, y, _ in upd) * 1e-3
    upd_flops = sum(f for _, _, f in upd)
    info = ch.info()

    # verification of the last timed factorization (outside the timed region)
    A0 = TileMatrix(desc, rank)
    A0.buf.copy_(pristine)
    res = ch.residual(A0)
    del A0

    # end to end through the public API with HOST buffers (pinned): H2D of the step's tiles,
    # factorization, D2H of the factor — all inside the timed region
    e2e = None
    if not a.no_e2e:
        nbytes = M.buf.numel() * 8
        hin = torch.empty(M.buf.shape, dtype=torch.float64).pin_memory()
        hin.copy_(pristine)
        hout = torch.empty(M.buf.shape, dtype=torch.float64).pin_memory()
        ch.factor_from_host(hin, hout)          # warm-up (also faults the pinned pages in)
        barrier()
        ke = max(1, min(a.steps, 3))
        t0 = time.perf_counter()
        for _ in range(ke):
            ch.factor_from_host(hin, hout)
            torch.cuda.current_stream().synchronize()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        same = bool(torch.equal(hout, M.buf.cpu())) if N <= 16384 else None
        e2e = {"value": flops * ke / dt / 1e12, "unit": UNIT, "h2d_bytes_per_step": nbytes * world,
               "d2h_bytes_per_step": nbytes * world, "steps": ke, "ms_per_step": dt / ke * 1e3,
               "api": "TiledCholesky.factor_from_host(pinned tiles) -> pinned factor", "info": ch.info()}
        if same is not None:
            e2e["matches_device_path"] = same
        del hin, hout
    del pristine

    # FP64 roofline denominator: MEASURED_PEAKS.json has no FP64 entry, so measure it here
    peak = None
    if rank == 0:
        import ctypes
        v = ctypes.c_double()
        st = torch.cuda.current_stream().cuda_stream
        peaks = []
        for _ in range(3):
            _lib.call("chol_fp64_peak", 1, 20000, ctypes.byref(v), st)
            peaks.append(v.value / 1e12)
        peak = max(peaks)

    if rank != 0:
        runtime.finalize()
        return 0
    tf = flops * a.steps / elapsed / 1e12
    achieved = upd_flops / upd_s / 1e12 if upd_s > 0 else None
    traffic, traffic_note = None, None
    ncu_json = os.path.join(ROOT, "profiles", "update_kernel_ncu.json")   # one `ncu --set full` capture, see profiles/
    if os.path.exists(ncu_json):
        with open(ncu_json) as f:
            cap = json.load(f)
        traffic = cap["dram_bytes_read"] + cap["dram_bytes_write"]
        traffic_note = cap["note"]
    line = {"metric": METRIC, "value": tf, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": elapsed / a.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(a, g.P, g.Q),
            "backward_error": res["fro"], "residual_inf": res["inf"], "info": info,
            "frac_of_fp64_peak": tf / (peak * world) if peak else None,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "gemm_nt_dmma_kernel (fused SYRK+GEMM trailing update, FP64 DMMA)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if (achieved and peak) else None,
                         "peak_source": "measured live: chol_fp64_peak (DMMA m8n8k4 chains, all SMs, burst); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "launches_timed": len(upd), "share_of_step": upd_s / elapsed if elapsed else None,
                         "traffic": traffic, "traffic_note": traffic_note}}
    if not a.no_cpu_baseline and world == 1:
        try:
            cores = host_cores()
            r = run_cpu_reference_once(cores)
            line["cpu_baseline"] = {"value": r["tflops"], "unit": UNIT, "cores": cores, "kind": r["kind"],
                                    "sample": r["sample"], "tiled_single_worker": run_cpu_tiled_single_worker()}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "port",
                                    "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    runtime.finalize()
    return 0


if __name__ == "__main__":
    args = parse_args()
    sys.exit(main_reference(args) if args.impl == "reference" else main_ours(args))
