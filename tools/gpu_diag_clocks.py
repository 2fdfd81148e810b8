"""Phase timestamps of the 128x128 diagonal-block kernel (debug build with -DCHOL_DIAG_CLOCKS, loaded through
CHOL_LIB_PATH).  Development tool."""
import ctypes as C, os, sys
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
raw = C.CDLL(_lib.LIB_PATH)
for rep in range(4):
    S2 = S.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call("chol_potrf_tile", b, S2.data_ptr(), b, work.data_ptr(), info.data_ptr(), 0, st)
    e1.record()
    torch.cuda.synchronize()
    clk = (C.c_longlong * 16)()
    raw.chol_debug_diag_clocks(clk)
    c = list(clk)
    names = ["load", "potrf0", "subst0", "upd0", "potrf1", "subst1", "upd1", "potrf2", "subst2", "upd2", "potrf3", "inv3(end loop)"]
    d = [c[1] - c[0]] + [c[i + 1] - c[i] for i in range(1, 11)] + [c[14] - c[11], c[15] - c[14]]
    names.append("store")
    print(f"rep {rep}: {e0.elapsed_time(e1)*1e3:.1f} us total; cycles " + " ".join(f"{n}={v}" for n, v in zip(names, d)), flush=True)

# ---- batched kernel: phases of one mid-grid CTA (under the contention of the full grid)
n, batch = 256, 10000
A0 = torch.empty(batch, n, n, dtype=torch.float64, device=dev)
for i in range(batch):
    _lib.call("chol_plgsy_tile", float(n), n, n, A0[i].data_ptr(), n, n, 0, 0, n, 42 + i, st)
binfo = torch.zeros(batch, dtype=torch.int32, device=dev)
for rep in range(3):
    Ab = A0.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call("chol_potrf_batched", n, batch, Ab.data_ptr(), n, n * n, binfo.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    clk = (C.c_longlong * 32)()
    raw.chol_debug_batched_clocks(clk)
    c = list(clk)
    parts = []
    for j in range(8):
        parts.append(f"j{j}: upd={c[1+3*j]-c[3*j]} potrf+pass1={c[2+3*j]-c[1+3*j]} trsm={c[3+3*j]-c[2+3*j]}")
    print(f"batched rep {rep}: {e0.elapsed_time(e1):.3f} ms; CTA total {c[24]-c[0]} cycles; " + " | ".join(parts), flush=True)
