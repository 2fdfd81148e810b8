"""Whole-matrix factorization check + timing on one GPU (development tool, run under gpurun).
usage: python tools/gpu_factor.py [N b]..."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dense_linear_app_b200 import _lib  # noqa: E402
from dense_linear_app_b200.cholesky import TiledCholesky  # noqa: E402
from dense_linear_app_b200.tiles import TileDesc, TileMatrix  # noqa: E402

_lib.call("chol_init", 0)
args = [int(x) for x in sys.argv[1:]] or [1000, 128, 4096, 512, 16384, 1024]
for N, b in zip(args[::2], args[1::2]):
    A = TileMatrix(TileDesc.square(N, b)).generate(float(N), 42)
    torch.cuda.synchronize()
    A0 = A.clone() if N <= 32768 else None
    t0 = time.time()
    ch = TiledCholesky(A)
    torch.cuda.synchronize()
    t_plan = time.time() - t0
    for rep in range(2 if N <= 32768 else 1):
        if rep:
            A.buf.copy_(A0.buf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.time()
        e0.record()
        ch.factor()
        e1.record()
        t_enq = time.time() - t0
        torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) * 1e-3
        print(f"N={N} b={b} rep={rep} plan={t_plan*1e3:.1f}ms enqueue={t_enq*1e3:.1f}ms time={dt*1e3:.2f}ms "
              f"{N**3/3/dt/1e12:.2f} TFLOP/s info={ch.info()}", flush=True)
    if os.environ.get("CHOL_FAST_ONLY"):
        del A, A0, ch
        torch.cuda.empty_cache()
        continue
    if N <= 8192:
        from scipy.linalg import lapack
        full = A0.to_numpy()
        full = np.tril(full) + np.tril(full, -1).T
        Lr, info = lapack.dpotrf(full, lower=1, clean=1)
        Lg = np.tril(A.to_numpy())
        print("   max|L-Lref|/max|L| =", np.abs(Lg - Lr).max() / np.abs(Lr).max(),
              " bwd err (host) =", np.linalg.norm(Lg @ Lg.T - full) / np.linalg.norm(full), flush=True)
    if A0 is None:
        A0 = TileMatrix(TileDesc.square(N, b)).generate(float(N), 42)
    t0 = time.time()
    r = ch.residual(A0)
    torch.cuda.synchronize()
    print(f"   residual: {r}  ({time.time()-t0:.2f}s)", flush=True)
    del A, A0, ch
    torch.cuda.empty_cache()
