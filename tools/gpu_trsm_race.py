"""Repeat the panel TRSM (and POTRF) many times on identical inputs, with and without a concurrent update
kernel on another stream, and report any run whose result differs from the first (development tool)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dense_linear_app_b200 import _lib
_lib.call("chol_init", 0)
lib = _lib.load()
dev = torch.device("cuda:0")
b, m = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 31
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
torch.manual_seed(1)
A = torch.rand(b, b, dtype=torch.float64, device=dev)
S = (A @ A.T + b * torch.eye(b, dtype=torch.float64, device=dev)).contiguous()
work = torch.empty(lib.chol_potrf_tile_workspace(b) // 8, dtype=torch.float64, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
tiles0 = torch.rand(m, b, b, dtype=torch.float64, device=dev)
tiles = tiles0.clone()
ptrs = torch.tensor([tiles[i].data_ptr() for i in range(m)], dtype=torch.int64, device=dev)
# background load: a trailing update on another stream
nt = 12
pan = torch.rand(nt, b, b, dtype=torch.float64, device=dev)
C = torch.zeros(nt * (nt + 1) // 2, b, b, dtype=torch.float64, device=dev)
tasks = []; idx = 0
for i in range(nt):
    for j in range(i + 1):
        tasks.append([C[idx].data_ptr(), pan[i].data_ptr(), pan[j].data_ptr(), int(i == j)]); idx += 1
dt = torch.tensor(tasks, dtype=torch.int64, device=dev)
bg = torch.cuda.Stream()
hi = torch.cuda.Stream(priority=-1)
ref_L = ref_X = None
for load in (0, 1):
    bad = 0
    for rep in range(reps):
        S2 = S.clone(); tiles.copy_(tiles0)
        torch.cuda.synchronize()
        if load:
            with torch.cuda.stream(bg):
                for _ in range(3):
                    _lib.call("chol_gemm_tasks", dt.data_ptr(), len(tasks), b, b, b, b, b, b, -1.0, 1.0, bg.cuda_stream)
        with torch.cuda.stream(hi):
            _lib.call("chol_potrf_tile", b, S2.data_ptr(), b, work.data_ptr(), info.data_ptr(), 0, hi.cuda_stream)
            _lib.call("chol_trsm_tiles", b, S2.data_ptr(), b, work.data_ptr(), ptrs.data_ptr(), m, b, None, hi.cuda_stream)
        torch.cuda.synchronize()
        if ref_L is None:
            ref_L, ref_X = S2.clone(), tiles.clone()
            continue
        if not torch.equal(torch.triu(S2), torch.triu(ref_L)):
            bad += 1; print(f"load={load} rep={rep}: POTRF result differs, max {float((S2-ref_L).abs().max()):.3e}")
        d = (tiles - ref_X).abs()
        if float(d.max()) != 0.0:
            bad += 1
            nz = (d > 0).nonzero()
            t0 = int(nz[0, 0])
            sub = (d[t0] > 0).nonzero()
            print(f"load={load} rep={rep}: TRSM differs in {nz.shape[0]} elements, max {float(d.max()):.3e}; tile {t0}: cols {int(sub[:,0].min())}..{int(sub[:,0].max())} rows {int(sub[:,1].min())}..{int(sub[:,1].max())}", flush=True)
    print(f"load={load}: {bad} of {reps} runs differ from the first", flush=True)
