"""The opt-in SM partition (CHOL_PANEL_SMS=16, CUDA green contexts) against the ordinary schedule: same bits, timing of
both.  Kept out of the pytest suite on purpose: a process that creates green contexts is killed by Nsight Compute
2025.2, and the GPU tests must stay profilable.  usage: python tools/gpu_partition_check.py [N b]..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import hashlib, os, sys, torch
sys.path.insert(0, %r)
from dense_linear_app_b200 import _lib
from dense_linear_app_b200.cholesky import TiledCholesky
from dense_linear_app_b200.tiles import TileDesc, TileMatrix
_lib.call("chol_init", 0)
N, b = int(sys.argv[1]), int(sys.argv[2])
A = TileMatrix(TileDesc.square(N, b)).generate(float(N), 42)
A0 = A.clone()
ch = TiledCholesky(A)
best = 1e9
for rep in range(3):
    A.buf.copy_(A0.buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); ch.factor(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(hashlib.sha256(A.buf.cpu().numpy().tobytes()).hexdigest(), best, ch.info(), ch.s_potrf is not None)
''' % ROOT
args = [int(x) for x in sys.argv[1:]] or [2048, 256, 4096, 512, 16384, 1024]
ok = True
for N, b in zip(args[::2], args[1::2]):
    out = {}
    for sms in ("0", "16"):
        env = dict(os.environ, CHOL_PANEL_SMS=sms)
        r = subprocess.run([sys.executable, "-c", CHILD, str(N), str(b)], env=env, capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            print(f"N={N} b={b} CHOL_PANEL_SMS={sms}: exit {r.returncode}: {r.stderr[-300:]}")
            ok = False
            continue
        sha, ms, info, part = r.stdout.split()[-4:]
        out[sms] = (sha, float(ms), int(info), part)
    if len(out) == 2:
        same = out["0"][0] == out["16"][0]
        ok = ok and same and out["16"][3] == "True"
        print(f"N={N} b={b}: two-stream {out['0'][1]:.2f} ms, partition {out['16'][1]:.2f} ms, info {out['0'][2]}/{out['16'][2]}, "
              f"partition active: {out['16'][3]}, bit-identical: {same}", flush=True)
sys.exit(0 if ok else 1)
