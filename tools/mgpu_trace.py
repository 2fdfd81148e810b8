"""CUDA-event timeline of one factorization on every rank (development tool): per step the start/end
of POTRF, TRSM, the arrival of L_kk / the panel, and the three stages of the trailing update.
    python -m torch.distributed.run --nproc-per-node N ... tools/mgpu_trace.py [N] [tile]
    python tools/mgpu_trace.py 16384 1024            (one GPU)
Writes gpurun_out/trace_<world>gpu_N<N>_rank<r>.json = [[name, k, ms], ...]."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dense_linear_app_b200 import runtime  # noqa: E402
from dense_linear_app_b200.cholesky import TiledCholesky  # noqa: E402
from dense_linear_app_b200.grid import ProcessGrid  # noqa: E402
from dense_linear_app_b200.tiles import TileDesc, TileMatrix  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
rank, world = runtime.init()
g = ProcessGrid.for_world(world)
desc = TileDesc(b, b, b * b, N, N, 0, 0, N, N, g.P, g.Q)
M = TileMatrix(desc, rank).generate(float(N), 42)
pristine = M.buf.clone()
ch = TiledCholesky(M)
for rep in range(3):
    M.buf.copy_(pristine)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()
    if rep == 2:
        ch.trace = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if rep == 2:
        ch.trace.append(("start", -1, e0))
    ch.factor()
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"rep {rep}: {e0.elapsed_time(e1):.2f} ms  {N ** 3 / 3 / e0.elapsed_time(e1) / 1e9:.2f} TFLOP/s", flush=True)
tr = ch.trace_ms()
ch.trace = None
assert ch.info() == 0
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"trace_{world}gpu_N{N}_rank{rank}.json"), "w") as f:
    json.dump(tr, f)
ch.close()
runtime.finalize()
