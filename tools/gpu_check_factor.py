"""Factor on one GPU, compare with scipy dpotrf element-wise per tile, print device residual (development tool)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dense_linear_app_b200 import _lib
from dense_linear_app_b200.cholesky import TiledCholesky
from dense_linear_app_b200.tiles import TileDesc, TileMatrix
from scipy.linalg import lapack
_lib.call("chol_init", 0)
for N, b in [(4096, 512), (8192, 1024), (12288, 1024)]:
    M = TileMatrix(TileDesc.square(N, b)).generate(float(N), 42)
    M0 = M.clone()
    A = M.to_numpy(); A = np.tril(A) + np.tril(A, -1).T
    ch = TiledCholesky(M); ch.factor(); info = ch.info()
    L = np.tril(M.to_numpy())
    Lref, _ = lapack.dpotrf(A, lower=1, clean=1)
    nt = N // b
    worst = []
    for i in range(nt):
        for j in range(i + 1):
            d = np.abs(L[i*b:(i+1)*b, j*b:(j+1)*b] - Lref[i*b:(i+1)*b, j*b:(j+1)*b]).max()
            worst.append((d, i, j))
    worst.sort(reverse=True)
    res = ch.residual(M0)
    hb = np.linalg.norm(L @ L.T - A) / np.linalg.norm(A) if N <= 8192 else float("nan")
    print(f"N={N} b={b} info={info} max|dL|/max|L|={worst[0][0]/np.abs(Lref).max():.2e} worst tiles={[(i,j,f'{d:.1e}') for d,i,j in worst[:4]]} host_bwd={hb:.2e} device_res={res['fro']:.2e}", flush=True)
