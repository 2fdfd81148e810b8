"""BASELINE configs[4]: 10 000 independent SPD matrices, n=256, FP64, one B200 (the ArmoniK
many-task workload).  Times chol_potrf_batched with CUDA events, checks info == 0 everywhere and
a sample of the factors against LAPACK dpotrf.  Prints one JSON line (development measurement,
not the headline bench).   usage: python tools/bench_batched.py [batch] [n]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dense_linear_app_b200 import _lib, tile_ops  # noqa: E402
from scipy.linalg import lapack  # noqa: E402  (sample check against LAPACK dpotrf; tools never touch oracle/)

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
_lib.call("chol_init", 0)
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
A = torch.empty(batch, n, n, dtype=torch.float64, device=dev)
for i in range(batch):   # same generator as the tiled path: bump = n, seed = 42 + matrix index
    _lib.call("chol_plgsy_tile", float(n), n, n, A[i].data_ptr(), n, n, 0, 0, n, 42 + i, st)
A0 = A.clone()
times = []
for rep in range(4):
    A.copy_(A0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    info = tile_ops.potrf_batched(A)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) * 1e-3)
t = min(times[1:])
bad = int((info != 0).sum().item())
errs = []
for i in (0, 1, batch // 2, batch - 1):
    full = A0[i].cpu().numpy().T
    ref, inf = lapack.dpotrf(np.tril(full) + np.tril(full, -1).T, lower=1, clean=1)
    assert inf == 0
    got = A[i].cpu().numpy().T
    errs.append(float(np.abs(np.tril(got) - np.tril(ref)).max() / np.abs(ref).max()))
alg_bytes = batch * 8.0 * n * (n + 1)
print(json.dumps({"workload": f"batched POTRF, {batch} x (n={n}) FP64 (BASELINE configs[4])", "seconds": t,
                  "tflops": batch * n ** 3 / 3 / t / 1e12, "matrices_per_s": batch / t,
                  "hbm_gbs_algorithmic": alg_bytes / t / 1e9, "nonzero_info": bad,
                  "max_rel_err_vs_lapack_sample": max(errs)}))
