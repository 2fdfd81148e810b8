"""One trailing-update launch (the fused SYRK+GEMM of a step with nt panel tiles, b=1024) a few times, for
ncu captures and timing (development tool).  usage: python tools/gpu_update_only.py [nt] [reps] [ctas_per_sm]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dense_linear_app_b200 import _lib
_lib.call("chol_init", 0)
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
b = 1024
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 15
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
occ = int(sys.argv[3]) if len(sys.argv) > 3 else 2
pan = torch.rand(nt, b, b, dtype=torch.float64, device=dev)
C = torch.rand(nt * (nt + 1) // 2, b, b, dtype=torch.float64, device=dev)
tasks = []; idx = 0
for i in range(nt):                      # row-major over (i, j), like the plan
    for j in range(i + 1):
        tasks.append([C[idx].data_ptr(), pan[i].data_ptr(), pan[j].data_ptr(), int(i == j)]); idx += 1
dt = torch.tensor(tasks, dtype=torch.int64, device=dev)
fl = sum(b ** 3 if t[3] else 2 * b ** 3 for t in tasks)
for rep in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call("chol_gemm_tasks_ex", dt.data_ptr(), len(tasks), b, b, b, b, b, b, -1.0, 1.0, occ, st)
    e1.record()
    torch.cuda.synchronize()
    print(f"nt={nt} tasks={len(tasks)} occ={occ} rep={rep}: {e0.elapsed_time(e1):.3f} ms  {fl / e0.elapsed_time(e1) / 1e9:.2f} TFLOP/s", flush=True)
