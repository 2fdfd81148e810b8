"""POTRF tile + panel TRSM (m tiles) once each at b=1024, for an ncu launch list (development tool)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dense_linear_app_b200 import _lib
_lib.call("chol_init", 0)
lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
b = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2
A = torch.rand(b, b, dtype=torch.float64, device=dev)
S = (A @ A.T + b * torch.eye(b, dtype=torch.float64, device=dev)).contiguous()
work = torch.empty(lib.chol_potrf_tile_workspace(b) // 8, dtype=torch.float64, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
tiles = torch.rand(m, b, b, dtype=torch.float64, device=dev)
ptrs = torch.tensor([tiles[i].data_ptr() for i in range(m)], dtype=torch.int64, device=dev)
for rep in range(3):
    S2 = S.clone()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    _lib.call("chol_potrf_tile", b, S2.data_ptr(), b, work.data_ptr(), info.data_ptr(), 0, st)
    e1.record()
    _lib.call("chol_trsm_tiles", b, S2.data_ptr(), b, work.data_ptr(), ptrs.data_ptr(), m, b, None, st)
    e2.record()
    torch.cuda.synchronize()
    print(f"rep {rep}: potrf {e0.elapsed_time(e1)*1e3:.0f} us  trsm({m}) {e1.elapsed_time(e2)*1e3:.0f} us  info {int(info.item())}")
