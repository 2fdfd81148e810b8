"""Batched POTRF (configs[4]) with the kernel chosen by CHOL_BATCHED_LL: time + a hash of all factors, so that two
runs (e.g. =4, without lookahead, and =8, with) can be compared bit for bit.  Development tool."""
import hashlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dense_linear_app_b200 import _lib
_lib.call("chol_init", 0)
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
for n, batch in ((256, 10000), (224, 64), (128, 100), (64, 37), (32, 5)):
    A0 = torch.empty(batch, n, n, dtype=torch.float64, device=dev)
    for i in range(batch):
        _lib.call("chol_plgsy_tile", float(n), n, n, A0[i].data_ptr(), n, n, 0, 0, n, 42 + i, st)
    if n == 224:
        A0[7, 100, 100] = -1.0          # a bad pivot
    info = torch.zeros(batch, dtype=torch.int32, device=dev)
    best = 1e9
    for rep in range(4):
        A = A0.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("chol_potrf_batched", n, batch, A.data_ptr(), n, n * n, info.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    h = hashlib.sha256(A.cpu().numpy().tobytes()).hexdigest()[:16]
    bad = int((info != 0).sum().item())
    print(f"mode={os.environ.get('CHOL_BATCHED_LL', 'default')} n={n} batch={batch}: {best:.3f} ms  "
          f"{batch * n**3 / 3 / best / 1e9:.2f} TFLOP/s  nonzero_info={bad} first={info[7].item() if n == 224 else 0} sha={h}", flush=True)
