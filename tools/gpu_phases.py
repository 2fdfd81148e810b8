"""Per-phase timings of one panel step (development tool): POTRF tile, TRSM panel, trailing update."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dense_linear_app_b200 import _lib
_lib.call("chol_init", 0)
lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream

def timeit(fn, reps=10, pre=None):
    for _ in range(2):
        if pre: pre()
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if pre: pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps

for b in (512, 1024, 2048):
    A = torch.rand(b, b, dtype=torch.float64, device=dev)
    S = (A @ A.T + b * torch.eye(b, dtype=torch.float64, device=dev)).contiguous()
    S2 = S.clone()
    work = torch.empty(lib.chol_potrf_tile_workspace(b) // 8, dtype=torch.float64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    tp = timeit(lambda: _lib.call("chol_potrf_tile", b, S2.data_ptr(), b, work.data_ptr(), info.data_ptr(), 0, st),
                pre=lambda: S2.copy_(S))
    # diag kernel alone: potrf on a 128 block
    D = S[:128, :128].contiguous(); D2 = D.clone()
    td = timeit(lambda: _lib.call("chol_potrf_tile", 128, D2.data_ptr(), 128, work.data_ptr(), info.data_ptr(), 0, st),
                pre=lambda: D2.copy_(D))
    print(f"b={b}: potrf_tile {tp*1e3:.0f} us ({b**3/3/tp/1e9:.2f} TF)   128-block diag kernel {td*1e3:.0f} us", flush=True)
    for m in (1, 7, 15, 31, 63):
        tiles = torch.rand(m, b, b, dtype=torch.float64, device=dev)
        ptrs = torch.tensor([tiles[i].data_ptr() for i in range(m)], dtype=torch.int64, device=dev)
        scratch = torch.empty(m * 8, dtype=torch.int64, device=dev)
        tt = timeit(lambda: _lib.call("chol_trsm_tiles", b, S2.data_ptr(), b, work.data_ptr(), ptrs.data_ptr(), m, b,
                                      scratch.data_ptr(), st), reps=5)
        print(f"    trsm panel m={m}: {tt*1e3:.0f} us ({m*b**3/tt/1e9:.2f} TF)", flush=True)
    if b == 1024:
        for nt in (2, 4, 8, 16):
            pan = torch.rand(nt, b, b, dtype=torch.float64, device=dev)
            C = torch.rand(nt * (nt + 1) // 2, b, b, dtype=torch.float64, device=dev)
            tasks = []; idx = 0
            for i in range(nt):
                for j in range(i + 1):
                    tasks.append([C[idx].data_ptr(), pan[i].data_ptr(), pan[j].data_ptr(), int(i == j)]); idx += 1
            dt = torch.tensor(tasks, dtype=torch.int64, device=dev)
            fl = sum(b**3 if t[3] else 2*b**3 for t in tasks)
            tu = timeit(lambda: _lib.call("chol_gemm_tasks", dt.data_ptr(), len(tasks), b, b, b, b, b, b, -1.0, 1.0, st), reps=5)
            print(f"    update nt={nt} ({len(tasks)} tasks): {tu*1e3:.0f} us ({fl/tu/1e9:.2f} TF)", flush=True)
