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

Other legs (each prints its own JSON line with roofline / cpu_baseline / e2e):
    --config c1        BASELINE configs[1]: N=16384, tile 1024, one B200
    --config batched   BASELINE configs[4]: 10 000 SPD matrices n=256 through chol_potrf_batched
At 1 GPU the headline line also carries short measurements of those two and of the tile-worker
path (blob -> blob through worker.execute) under "other_configs".
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
    ap.add_argument("--config", default="headline", choices=["headline", "c1", "batched"])
    ap.add_argument("--batch", type=int, default=10000)
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-lead-check", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lookahead", action="store_true")
    a = ap.parse_args()
    if a.config == "c1":
        a.N, a.tile = 16384, 1024
    return a


def workload_config(a, P, Q):
    which = {(65536, 1024): "configs[2]", (16384, 1024): "configs[1]", (131072, 2048): "configs[3]",
             (4096, 512): "configs[0]"}.get((a.N, a.tile), "custom size")
    return {"workload": f"FP64 random SPD N={a.N}, tile {a.tile}, lower Cholesky A=LL^T (BASELINE {which})",
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


def run_cpu_monolithic(N: int, threads: int) -> dict:
    """monolithic LAPACK dpotrf('L') on all host cores through the oracle's
    OpenBLAS port (the call lapacke_dpotrf.c:54 makes), on the same generator as the GPU runs."""
    from oracle import oracle as O
    A = O.plgsy(float(N), N, 42)
    t0 = time.time()
    _, info = O.blas_potrf(A, threads=threads)
    dt = time.time() - t0
    return {"workload": f"monolithic OpenBLAS dpotrf N={N}, {threads} threads (oracle port)", "seconds": dt,
            "tflops": N ** 3 / 3 / dt / 1e12, "info": int(info)}


def run_cpu_batched(n: int, count: int) -> dict:
    """CPU leg of configs[4]: a loop of LAPACK dpotrf over `count` n x n matrices, one BLAS thread
    (one POTRF task per worker, worker_distrib.cpp:238; benchmark.c:173-175 forces 1 BLAS thread)."""
    from oracle import oracle as O
    mats = [O.plgsy(float(n), n, 42 + i) for i in range(count)]
    t0 = time.time()
    for m in mats:
        _, info = O.blas_potrf(m, threads=1)
        assert info == 0
    dt = time.time() - t0
    return {"seconds": dt, "tflops": count * n ** 3 / 3 / dt / 1e12, "matrices_per_s": count / dt}


def reference_config(a, P, Q, sample: dict) -> dict:
    """The arm's own config (so the two arms name the same workload) with what this arm really times
    spelled out: the reference program's size is fixed in its source."""
    c = workload_config(a, P, Q)
    c["workload"] += (f" — reference arm: each step is a BOUNDED SAMPLE of it, one factorization at "
                      f"N={sample['N']} ({sample['kind']})")
    c["N_timed"] = sample["N"]
    c["tile_timed"] = None
    return c


def main_reference(a) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from dense_linear_app_b200.grid import ProcessGrid
    g = ProcessGrid.for_world(a.gpus)
    cores = host_cores()
    if a.config == "batched":
        count = min(a.batch, 2000)
        for _ in range(min(a.warmup, 1)):
            run_cpu_batched(a.n, 200)
        runs = [run_cpu_batched(a.n, count) for _ in range(max(1, min(a.steps, 3)))]
        tf = sum(count * a.n ** 3 / 3 for _ in runs) / sum(r["seconds"] for r in runs) / 1e12
        ms = sum(r["seconds"] for r in runs) / len(runs) * 1e3
        sample = f"loop of OpenBLAS dpotrf over {count} of the {a.batch} matrices (n={a.n}), 1 thread, per step"
        line = {"impl": "reference", "metric": "fp64_batched_cholesky_tflops", "value": tf, "unit": UNIT,
                "n_gpus": a.gpus, "steps": len(runs), "warmup": min(a.warmup, 1), "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": batched_config(a),
                "cpu_baseline": {"value": tf, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
                "e2e": {"value": tf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0
    for _ in range(a.warmup):
        run_cpu_reference_once(cores)
    runs = [run_cpu_reference_once(cores) for _ in range(a.steps)]
    tf = sum(r["N"] ** 3 / 3 for r in runs) / sum(r["seconds"] for r in runs) / 1e12
    ms = sum(r["seconds"] for r in runs) / len(runs) * 1e3
    line = {"impl": "reference", "metric": METRIC, "value": tf, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": reference_config(a, g.P, g.Q, runs[0]),
            "cpu_baseline": {"value": tf, "unit": UNIT, "cores": cores, "kind": runs[0]["kind"],
                             "sample": runs[0]["sample"] + f" per step, {a.steps} steps"},
            "tiled_single_worker": run_cpu_tiled_single_worker(),
            "monolithic_N16384_all_cores": run_cpu_monolithic(16384, cores),
            "e2e": {"value": tf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def batched_config(a) -> dict:
    return {"workload": f"batched small Cholesky: {a.batch} SPD matrices n={a.n} FP64 on 1 B200 (BASELINE configs[4])",
            "batch": a.batch, "n": a.n, "generator": "dplgsy-style LCG, bump=n, seed=42+index",
            "flops": "batch*n^3/3", "bytes": "8*n*(n+1) per matrix (lower triangle in + out)",
            "l2": "inputs larger than L2 (%.2f GB >> 126 MB)" % (a.batch * a.n * a.n * 8 / 1e9)}


# ---- checks and side measurements of our arm ----------------------------------------------------------
def lead_block_check(M, desc, rank, world, dist, m: int) -> dict:
    """chol(A)[:m,:m] == chol(A[:m,:m]).  The leading m x m block of the factor (gathered
    from its owners) against LAPACK dpotrf of the leading block of the same matrix on the host cores
    (scipy's OpenBLAS — the library family the reference calls).  Element gate of north_star:
    |L - Lref| <= 1e-10 * max(|Lref|, 1e-3 max|Lref|)."""
    import numpy as np
    import torch
    from dense_linear_app_b200.tiles import TileDesc, TileMatrix
    b, N = desc.mb, desc.m
    nb = m // b
    m = nb * b
    dev = M.device
    Lg = torch.zeros((m, m), dtype=torch.float64, device=dev)      # [col, row] like a col-major matrix
    for i, j in M.layout.tiles():
        if i < nb and j < nb:
            Lg[j * b:(j + 1) * b, i * b:(i + 1) * b] = M.tile(i, j)
    if world > 1:
        dist.reduce(Lg, dst=0)
    if rank != 0:
        return None
    t0 = time.time()
    d1 = TileDesc(b, b, b * b, desc.lm, desc.ln, 0, 0, m, m, 1, 1)        # same generator stream: bigM = lm
    A = TileMatrix(d1, 0, dev).generate(float(N), 42).to_numpy()
    A = np.tril(A) + np.tril(A, -1).T
    from scipy.linalg import lapack
    Lref, info = lapack.dpotrf(A, lower=1, clean=1)
    got = np.tril(Lg.cpu().numpy().T)
    scale = float(np.abs(Lref).max())
    diff = np.abs(got - Lref)
    gate = 1e-10 * np.maximum(np.abs(Lref), 1e-3 * scale)
    return {"m": m, "max_rel": float(diff.max() / scale), "element_gate_1e-10": bool((diff <= gate).all()),
            "cpu_info": int(info), "against": "scipy OpenBLAS dpotrf of A[:m,:m] on the host", "seconds": time.time() - t0}


def hbm_peak() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


def measure_batched(batch: int, n: int, steps: int, warmup: int, with_e2e: bool) -> dict:
    """configs[4] through tile_ops.potrf_batched: kernel time by CUDA events around the launch only
    (the restore from a pristine copy sits between the events of consecutive steps), info of every
    matrix, backward error of every matrix (torch.bmm, the checker)."""
    import torch
    from dense_linear_app_b200 import _lib, tile_ops
    dev = torch.device("cuda", torch.cuda.current_device())
    st = torch.cuda.current_stream().cuda_stream
    A0 = torch.empty(batch, n, n, dtype=torch.float64, device=dev)
    for i in range(batch):
        _lib.call("chol_plgsy_tile", float(n), n, n, A0[i].data_ptr(), n, n, 0, 0, n, 42 + i, st)
    A = torch.empty_like(A0)
    lib = _lib.load()
    ms = []
    n0 = 0
    for it in range(warmup + steps):
        A.copy_(A0)
        if it == warmup:
            n0 = lib.chol_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        info = tile_ops.potrf_batched(A)
        e1.record()
        torch.cuda.synchronize()
        if it >= warmup:
            ms.append(e0.elapsed_time(e1))
    launches = lib.chol_launch_count() - n0
    t = sum(ms) / len(ms) * 1e-3
    flops = batch * float(n) ** 3 / 3
    alg_bytes = batch * 8.0 * n * (n + 1)
    Lt = torch.triu(A)
    full = torch.triu(A0) + torch.triu(A0, 1).transpose(1, 2)
    err = float((torch.linalg.matrix_norm(full - Lt.transpose(1, 2) @ Lt) / torch.linalg.matrix_norm(full)).max().item())
    del Lt, full
    peak, src = hbm_peak()
    traffic, traffic_note = None, None
    ncu_json = os.path.join(ROOT, "profiles", "batched_kernel_ncu.json")     # one `ncu --set full` capture, see profiles/
    if os.path.exists(ncu_json) and batch == 10000 and n == 256:
        with open(ncu_json) as f:
            cap = json.load(f)
        traffic, traffic_note = cap["dram_bytes_read"] + cap["dram_bytes_write"], cap["note"]
    out = {"value": flops / t / 1e12, "unit": UNIT, "ms_per_step": t * 1e3, "ms_best": min(ms), "steps": steps,
           "matrices_per_s": batch / t, "nonzero_info": int((info != 0).sum().item()), "max_backward_error": err,
           "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "kernel": "potrf_batched_np_kernel (left-looking, DMMA, one CTA per matrix, 4 CTAs/SM)",
                        "achieved": alg_bytes / t / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg_bytes / t / 1e9 / peak, "peak_source": src, "traffic": traffic, "traffic_note": traffic_note,
                        "algorithmic_bytes": alg_bytes, "also_fp64": "flops / t vs the FP64 DMMA peak: see frac_of_fp64_peak"}}
    if with_e2e:
        hin = torch.empty(A0.shape, dtype=torch.float64).pin_memory()
        hin.copy_(A0)
        hout = torch.empty(A0.shape, dtype=torch.float64).pin_memory()
        hinfo = torch.empty(batch, dtype=torch.int32).pin_memory()

        def one():
            tile_ops.potrf_batched_from_host(hin, hout, hinfo)
            torch.cuda.current_stream().synchronize()
        one()
        ke = max(1, min(steps, 5))
        t0 = time.perf_counter()
        for _ in range(ke):
            one()
        dt = (time.perf_counter() - t0) / ke
        out["e2e"] = {"value": flops / dt / 1e12, "unit": UNIT, "h2d_bytes_per_step": A.numel() * 8,
                      "d2h_bytes_per_step": A.numel() * 8 + batch * 4, "steps": ke, "ms_per_step": dt * 1e3,
                      "api": "tile_ops.potrf_batched_from_host(pinned matrices) -> pinned factors + info, 16 pieces "
                             "pipelined over upload / factor / download streams",
                      "matches_device_path": bool(torch.equal(hout.to(dev), A)),
                      "nonzero_info": int((hinfo != 0).sum().item())}
        del hin, hout
    del A, A0
    torch.cuda.empty_cache()
    return out


def measure_worker_path(B: int, reps: int = 6) -> dict:
    """Entry point #2 (worker_distrib.cpp:212-262): one GEMM and one POTRF tile task, blob -> blob
    through worker.execute (JSON payload, host blobs, H2D, kernel, D2H, blob)."""
    import numpy as np
    from dense_linear_app_b200 import worker
    rng = np.random.default_rng(1)
    g = rng.standard_normal((B, B))
    spd = (g @ g.T + B * np.eye(B)).tobytes(order="F")
    blobs = {"c": rng.standard_normal((B, B)).tobytes(), "a": rng.standard_normal((B, B)).tobytes(),
             "b": rng.standard_normal((B, B)).tobytes(), "d": spd}
    pay_g = json.dumps({"op": "GEMM", "B": B, "inC": "c", "inAi": "a", "inAj": "b"})
    pay_p = json.dumps({"op": "POTRF", "B": B, "in": "d"})
    out = {}
    for name, pay, nblob in (("gemm", pay_g, 4), ("potrf", pay_p, 2)):
        worker.execute(pay, blobs)
        t0 = time.perf_counter()
        for _ in range(reps):
            worker.execute(pay, blobs)
        dt = (time.perf_counter() - t0) / reps
        out[name] = {"tasks_per_s": 1.0 / dt, "ms_per_task": dt * 1e3, "blob_GBps": nblob * B * B * 8 / dt / 1e9}
    return out


def other_configs(a) -> dict:
    """Short measurements of BASELINE configs[1] and configs[4] and of the tile-worker path, carried in
    the headline line at 1 GPU (their own legs: --config c1 / --config batched)."""
    import torch
    from dense_linear_app_b200.cholesky import TiledCholesky
    from dense_linear_app_b200.tiles import TileDesc, TileMatrix
    out = {}
    try:
        N, b = 16384, 1024
        M = TileMatrix(TileDesc.square(N, b)).generate(float(N), 42)
        pristine = M.buf.clone()
        ch = TiledCholesky(M)
        ms = []
        for it in range(6):
            M.buf.copy_(pristine)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ch.factor()
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ms.append(e0.elapsed_time(e1))
        A0 = TileMatrix(TileDesc.square(N, b))
        A0.buf.copy_(pristine)
        res = ch.residual(A0)
        t = sum(ms) / len(ms) * 1e-3
        out["configs[1] N=16384 tile 1024"] = {"value": N ** 3 / 3 / t / 1e12, "unit": UNIT, "ms_per_step": t * 1e3,
                                               "steps": len(ms), "info": ch.info(), "backward_error": res["fro"],
                                               "timed": "factorization only (restore outside the events)"}
        del M, A0, pristine, ch
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["configs[1] N=16384 tile 1024"] = {"error": repr(e)}
    try:
        out["configs[4] batched 10000 x 256"] = measure_batched(10000, 256, 5, 3, with_e2e=False)
    except Exception as e:  # noqa: BLE001
        out["configs[4] batched 10000 x 256"] = {"error": repr(e)}
    try:
        out["worker_path"] = {f"B={B}": measure_worker_path(B) for B in (512, 1024)}
    except Exception as e:  # noqa: BLE001
        out["worker_path"] = {"error": repr(e)}
    return out


def main_batched(a) -> int:
    import torch
    from dense_linear_app_b200 import _lib, runtime
    rank, world = runtime.init()
    # the path shards by matrices with no exchange: N GPUs = N independent replicas of the batch
    dev = torch.device("cuda", torch.cuda.current_device())
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    r = measure_batched(a.batch, a.n, a.steps, max(a.warmup, 3), with_e2e=not a.no_e2e)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([r["ms_per_step"]], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
    peak = None
    if rank == 0:
        import ctypes
        v = ctypes.c_double()
        _lib.call("chol_fp64_peak", 1, 20000, ctypes.byref(v), torch.cuda.current_stream().cuda_stream)
        peak = v.value / 1e12
    if rank != 0:
        runtime.finalize()
        return 0
    flops = a.batch * float(a.n) ** 3 / 3 * world
    line = {"metric": "fp64_batched_cholesky_tflops", "value": flops / (ms * 1e-3) / 1e12, "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": batched_config(a),
            "matrices_per_s": a.batch * world / (ms * 1e-3), "nonzero_info": r["nonzero_info"],
            "max_backward_error": r["max_backward_error"], "frac_of_fp64_peak": r["value"] / peak if peak else None,
            "clocks": clocks, "e2e": r.get("e2e"), "gpu_launches": r["gpu_launches"], "roofline": r["roofline"]}
    if not a.no_cpu_baseline and world == 1:
        c = run_cpu_batched(a.n, min(a.batch, 2000))
        line["cpu_baseline"] = {"value": c["tflops"], "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"loop of OpenBLAS dpotrf over {min(a.batch, 2000)} of the matrices, 1 thread"}
    print(json.dumps(line), flush=True)
    runtime.finalize()
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
    snapshot = M.buf.clone()                 # the device-path factor, to compare the e2e result with
    A0 = TileMatrix(desc, rank)
    A0.buf.copy_(pristine)
    res = ch.residual(A0)
    del A0
    lead = None
    if not a.no_lead_check:
        lead = lead_block_check(M, desc, rank, world, dist, m=min(8192, N))

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
        ke = max(1, a.steps)
        t0 = time.perf_counter()
        for _ in range(ke):
            ch.factor_from_host(hin, hout)
            torch.cuda.current_stream().synchronize()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        # the factor that came back to the host against the device-path factor of the timed steps
        same = True
        chunk = max(1, (1 << 28) // (b * b))
        for lo in range(0, hout.shape[0], chunk):
            same = same and bool(torch.equal(hout[lo:lo + chunk].to(dev), snapshot[lo:lo + chunk]))
        e2e = {"value": flops * ke / dt / 1e12, "unit": UNIT, "h2d_bytes_per_step": nbytes * world,
               "d2h_bytes_per_step": nbytes * world, "steps": ke, "ms_per_step": dt / ke * 1e3,
               "api": "TiledCholesky.factor_from_host(pinned tiles) -> pinned factor", "info": ch.info()}
        e2e["matches_device_path"] = same
        del hin, hout
    del pristine, snapshot

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
            "backward_error": res["fro"], "residual_inf": res["inf"], "info": info, "lead_block_check": lead,
            "lead_block_max_rel": lead["max_rel"] if lead else None,
            "frac_of_fp64_peak": tf / (peak * world) if peak else None,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "gemm_nt_dmma_kernel (fused SYRK+GEMM trailing update, FP64 DMMA)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if (achieved and peak) else None,
                         "peak_source": "measured live: chol_fp64_peak (DMMA m8n8k4 chains, all SMs, burst); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "launches_timed": len(upd), "share_of_step": upd_s / elapsed if elapsed else None,
                         "traffic": traffic, "traffic_note": traffic_note,
                         "traffic_config": "one ncu --set full capture of ONE 120-task update launch (15-tile panel, "
                                           "tile 1024; profiles/r02_update_kernel_ncu.md), per launch like "
                                           "`achieved`; not a capture of the benched run"}}
    if not a.no_cpu_baseline and world == 1:
        try:
            cores = host_cores()
            r = run_cpu_reference_once(cores)
            line["cpu_baseline"] = {"value": r["tflops"], "unit": UNIT, "cores": cores, "kind": r["kind"],
                                    "sample": r["sample"]}
            line["tiled_single_worker"] = run_cpu_tiled_single_worker()
            line["monolithic_N16384_all_cores"] = run_cpu_monolithic(16384, cores)
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "port",
                                    "sample": f"failed: {e}"}
    if world == 1 and a.config == "headline" and not a.no_other_configs:
        del M, ch
        torch.cuda.empty_cache()
        line["other_configs"] = other_configs(a)
    print(json.dumps(line), flush=True)
    runtime.finalize()
    return 0


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        sys.exit(main_reference(args))
    sys.exit(main_batched(args) if args.config == "batched" else main_ours(args))
