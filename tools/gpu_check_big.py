"""Locate residual errors at large N on one GPU: factor, device residual, worst residual tiles (development tool)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dense_linear_app_b200 import _lib
from dense_linear_app_b200.cholesky import TiledCholesky
from dense_linear_app_b200.tiles import TileDesc, TileMatrix
_lib.call("chol_init", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
M = TileMatrix(TileDesc.square(N, b)).generate(float(N), 42)
pristine = M.buf.clone()
ch = TiledCholesky(M)
for _ in range(reps):
    M.buf.copy_(pristine)
    ch.factor()
info = ch.info()
A0 = TileMatrix(TileDesc.square(N, b))
A0.buf.copy_(pristine)
res = ch.residual(A0)
worst = []
for i, j in A0.layout.tiles():
    t = A0.tile(i, j)
    if i == j:
        t = torch.triu(t)
    worst.append((float(t.abs().max().item()), i, j))
worst.sort(reverse=True)
print(f"N={N} b={b} info={info} residual={res['fro']:.3e} inf={res['inf']:.3e}  worst tiles (max|R|, i, j): {worst[:8]}", flush=True)
if worst[0][0] > 1e-6:
    d, i, j = worst[0]
    t = A0.tile(i, j)
    if i == j:
        t = torch.triu(t)
    idx = (t.abs() > 1e-6).nonzero()
    print(f"  tile ({i},{j}): {idx.shape[0]} elements > 1e-6; first (col,row): {idx[:12].tolist()}", flush=True)
