"""End-to-end time of TiledCholesky.factor_from_host (pinned host tiles -> factor -> pinned factor) next to the
device-resident factor(), one GPU.  Development tool.  usage: python tools/gpu_e2e.py [N b]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dense_linear_app_b200 import _lib
from dense_linear_app_b200.cholesky import TiledCholesky
from dense_linear_app_b200.tiles import TileDesc, TileMatrix
_lib.call("chol_init", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
M = TileMatrix(TileDesc.square(N, b)).generate(float(N), 42)
pristine = M.buf.clone()
ch = TiledCholesky(M)
for rep in range(2):
    M.buf.copy_(pristine)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); ch.factor(); e1.record(); torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1)
want = M.buf.clone()
hin = torch.empty(M.buf.shape, dtype=torch.float64).pin_memory(); hin.copy_(pristine)
hout = torch.empty(M.buf.shape, dtype=torch.float64).pin_memory()
del pristine
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ch.factor_from_host(hin, hout)
    torch.cuda.current_stream().synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    same = all(bool(torch.equal(hout[lo:lo + 256].cuda(), want[lo:lo + 256])) for lo in range(0, hout.shape[0], 256))
    print(f"N={N} b={b} lazy_steps={ch.lazy_steps}: device {dev_ms:.1f} ms = {N**3/3/dev_ms/1e9:.2f} TFLOP/s; "
          f"from host {dt:.1f} ms = {N**3/3/dt/1e9:.2f} TFLOP/s; identical={same} info={ch.info()}", flush=True)
