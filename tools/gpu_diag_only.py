"""Runs the 128x128 diagonal-block kernel a few times (target for an ncu source-level capture)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dense_linear_app_b200 import _lib
_lib.call("chol_init", 0)
lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
b = 128
A = torch.rand(b, b, dtype=torch.float64, device=dev)
S = (A @ A.T + b * torch.eye(b, dtype=torch.float64, device=dev)).contiguous()
work = torch.empty(lib.chol_potrf_tile_workspace(b) // 8, dtype=torch.float64, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
for _ in range(4):
    S2 = S.clone()
    _lib.call("chol_potrf_tile", b, S2.data_ptr(), b, work.data_ptr(), info.data_ptr(), 0, st)
torch.cuda.synchronize()
print("ok", int(info.item()))
