"""First-contact GPU probe (run under gpurun): FP64 peaks, per-op correctness vs scipy BLAS/LAPACK,
tile-op timings.  Writes gpurun_out/probe.json.  Development tool, not part of the product."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dense_linear_app_b200 import _lib  # noqa: E402

out = {}
lib = _lib.load()
_lib.call("chol_init", 0)
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream


def peak(kind, iters=20000):
    v = C.c_double()
    _lib.call("chol_fp64_peak", kind, iters, C.byref(v), st)
    return v.value / 1e12


out["peak_tflops"] = {"dfma": peak(0), "dmma_8w": peak(1), "dmma_4w": peak(2)}
print("peaks", out["peak_tflops"], flush=True)


def dev_cm(a):
    """numpy F-order (m,n) -> torch cuda tensor holding the same column-major bytes."""
    return torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)


def host_cm(t):
    return np.asfortranarray(t.cpu().numpy().T)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


from scipy.linalg import blas, lapack  # noqa: E402

rng = np.random.default_rng(0)
res = {}
for b in (4, 6, 64, 128, 200, 256, 512, 1024):
    Ai = np.asfortranarray(rng.uniform(-0.5, 0.5, (b, b)))
    Aj = np.asfortranarray(rng.uniform(-0.5, 0.5, (b, b)))
    Cm = np.asfortranarray(rng.uniform(-0.5, 0.5, (b, b)))
    r = {}
    # GEMM
    dC = dev_cm(Cm)
    dAi, dAj = dev_cm(Ai), dev_cm(Aj)
    _lib.call("chol_gemm_tile", b, dAi.data_ptr(), b, dAj.data_ptr(), b, dC.data_ptr(), b, st)
    ref = Cm - Ai @ Aj.T
    r["gemm"] = float(np.abs(host_cm(dC) - ref).max() / np.abs(ref).max())
    # SYRK
    dC = dev_cm(Cm)
    _lib.call("chol_syrk_tile", b, dAi.data_ptr(), b, dC.data_ptr(), b, st)
    full = Cm - Ai @ Ai.T
    ref = np.where(np.tril(np.ones((b, b), bool)), full, Cm)
    r["syrk"] = float(np.abs(host_cm(dC) - ref).max() / np.abs(ref).max())
    # POTRF
    S = Ai @ Ai.T + b * np.eye(b)
    S = np.asfortranarray(np.tril(S) + np.triu(np.full((b, b), 7.0), 1))  # strict upper = sentinel
    dS = dev_cm(S)
    wbytes = lib.chol_potrf_tile_workspace(b)
    work = torch.empty(max(wbytes // 8, 1), dtype=torch.float64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.call("chol_potrf_tile", b, dS.data_ptr(), b, work.data_ptr(), info.data_ptr(), 0, st)
    Lr, inf = lapack.dpotrf(S, lower=1, clean=0)
    got = host_cm(dS)
    r["potrf"] = float(np.abs(np.tril(got) - np.tril(Lr)).max() / np.abs(np.tril(Lr)).max())
    r["potrf_upper_untouched"] = bool(np.all(np.triu(got, 1) == np.triu(S, 1)))
    r["potrf_info"] = int(info.item())
    # TRSM
    Lk = np.asfortranarray(np.tril(Lr))
    dL = dev_cm(Lk)
    dA = dev_cm(Aj)
    _lib.call("chol_trsm_tile", b, dL.data_ptr(), b, dA.data_ptr(), b, work.data_ptr(), st)
    ref = blas.dtrsm(1.0, Lk, Aj, side=1, lower=1, trans_a=1, diag=0)
    r["trsm"] = float(np.abs(host_cm(dA) - ref).max() / np.abs(ref).max())
    # non-SPD -> info
    S2 = S.copy()
    k = b // 2
    S2[k, k] = -1.0
    dS2 = dev_cm(S2)
    info.zero_()
    _lib.call("chol_potrf_tile", b, dS2.data_ptr(), b, work.data_ptr(), info.data_ptr(), 0, st)
    _, inf2 = lapack.dpotrf(S2, lower=1, clean=0)
    r["potrf_info_bad"] = [int(info.item()), int(inf2)]
    res[b] = r
    print(b, r, flush=True)
out["ops"] = res

# timings
tm = {}
for b in (512, 1024, 2048, 4096):
    A = torch.rand(b, b, dtype=torch.float64, device=dev)
    B = torch.rand(b, b, dtype=torch.float64, device=dev)
    Cc = torch.rand(b, b, dtype=torch.float64, device=dev)
    t = timeit(lambda: _lib.call("chol_gemm_tile", b, A.data_ptr(), b, B.data_ptr(), b, Cc.data_ptr(), b, st))
    ts = timeit(lambda: _lib.call("chol_syrk_tile", b, A.data_ptr(), b, Cc.data_ptr(), b, st))
    tt = timeit(lambda: torch.matmul(A, B))
    S = (A @ A.T + b * torch.eye(b, dtype=torch.float64, device=dev)).contiguous()
    wbytes = lib.chol_potrf_tile_workspace(b)
    work = torch.empty(wbytes // 8, dtype=torch.float64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    S2 = S.clone()

    def potrf():
        S2.copy_(S)
        _lib.call("chol_potrf_tile", b, S2.data_ptr(), b, work.data_ptr(), info.data_ptr(), 0, st)

    tp = timeit(potrf)
    tcopy = timeit(lambda: S2.copy_(S))
    ttr = timeit(lambda: _lib.call("chol_trsm_tile", b, S2.data_ptr(), b, Cc.data_ptr(), b, work.data_ptr(), st))
    tm[b] = {"gemm_tflops": 2 * b ** 3 / t / 1e12, "syrk_tflops": b ** 3 / ts / 1e12,
             "cublas_dgemm_tflops": 2 * b ** 3 / tt / 1e12, "potrf_ms": (tp - tcopy) * 1e3, "trsm_ms": ttr * 1e3}
    print(b, tm[b], flush=True)
out["timing"] = tm

# grouped update: many tiles in one launch (the trailing-update shape)
b = 1024
nt = 12
tiles = torch.rand(nt * (nt + 1) // 2 + nt, b, b, dtype=torch.float64, device=dev)
tasks = []
base = tiles.data_ptr()
tsz = b * b * 8
pan = nt * (nt + 1) // 2
idx = 0
for i in range(nt):
    for j in range(i + 1):
        tasks.append([base + idx * tsz, base + (pan + i) * tsz, base + (pan + j) * tsz, 1 if i == j else 0])
        idx += 1
dt = torch.tensor(tasks, dtype=torch.int64, device=dev)
t = timeit(lambda: _lib.call("chol_gemm_tasks", dt.data_ptr(), len(tasks), b, b, b, b, b, b, -1.0, 1.0, st), reps=3)
flops = sum((b ** 3 if f else 2 * b ** 3) for *_, f in tasks)
out["grouped_update"] = {"ntasks": len(tasks), "tflops": flops / t / 1e12, "ms": t * 1e3}
print(out["grouped_update"], flush=True)
out["cublas_big_dgemm_tflops"] = None
A = torch.rand(8192, 8192, dtype=torch.float64, device=dev)
tt = timeit(lambda: torch.matmul(A, A), reps=3)
out["cublas_big_dgemm_tflops"] = 2 * 8192 ** 3 / tt / 1e12
print("cublas 8192 dgemm", out["cublas_big_dgemm_tflops"])

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/probe.json", "w") as f:
    json.dump(out, f, indent=1)
