"""``python -m dense_linear_app_b200.bench_sweep`` — the sweep harness (reference: benchmark.c).

For every N x NB x mapping x schedule x repeat it runs the driver as a CHILD PROCESS with the 16
positional arguments (benchmark.c:239-255), captures stdout, parses the ``Performance:`` and
``||A - LL^T||_inf / ||A||_inf =`` lines (benchmark.c:45-67) and appends one row to
``results/bench.csv`` with the reference's header (benchmark.c:114,282-285):

    timestamp,scheduler,mapping,ncpu,ngpu,N,NB,run_idx,ms,exit_code,gflops,rel_error

Defaults reproduce the reference grid (N in {1000,5000,8000,12000,16000}, NB in {128..512 step 64},
8 repeats with run 0 as the warm-up run — the reference's calibration run, benchmark.c:201).  The
StarPU scheduler axis becomes this framework's two schedules (``lookahead`` / ``inorder``) and the
mapping axis the GPU count (``1_b200`` ...).  The reference plot scripts read the CSV unchanged.
"""
from __future__ import annotations

import argparse
import os
import re
import subprocess
import sys
import time

HEADER = "timestamp,scheduler,mapping,ncpu,ngpu,N,NB,run_idx,ms,exit_code,gflops,rel_error\n"
NS = (1000, 5000, 8000, 12000, 16000)          # benchmark.c:76
NBS = (128, 192, 256, 320, 384, 448, 512)      # benchmark.c:80
SCHEDS = ("lookahead", "inorder")
REPEATS = 8                                    # benchmark.c:103


def parse_metrics(out: str) -> tuple[float, float]:
    """parse_metrics (benchmark.c:45-67): -1 when a line is missing."""
    g = re.search(r"Performance:\s*([-+0-9.eE]+|nan|inf) Gflop/s", out)
    r = re.search(r"\|\|A - LL\^T\|\|_inf / \|\|A\|\|_inf =\s*([-+0-9.eE]+|nan|inf)", out)
    return (float(g.group(1)) if g else -1.0), (float(r.group(1)) if r else -1.0)


def driver_argv(ncpu: int, ngpu: int, N: int, NB: int, p: int, q: int, seed: int) -> list[str]:
    """The 16 arguments benchmark.c:123-131,247-252 builds."""
    return [str(x) for x in (ncpu, ngpu, N, NB, NB, NB, NB * NB, N, N, 0, 0, N, N, p, q, seed)]


def run_in_process(sched: str, N: int, NB: int, seed: int = 42) -> dict:
    """One grid point on one GPU WITHOUT a child process: the driver's main() is called here and its stdout
    captured and parsed exactly like the child's (``--in-process``).  The reference always forks
    (benchmark.c:239-255); on a metered GPU box 560 interpreter + CUDA start-ups cost more than the sweep."""
    import contextlib
    import io
    from . import v6_test
    os.environ["STARPU_SCHED"] = sched
    os.environ["CHOL_LOOKAHEAD"] = "0" if sched == "inorder" else "1"
    buf = io.StringIO()
    t0 = time.time()
    try:
        with contextlib.redirect_stdout(buf):
            code = v6_test.main(["v6_test"] + driver_argv(0, 1, N, NB, 1, 1, seed))
    except Exception as e:                          # a crash of the child = non-zero exit, metrics missing
        sys.stderr.write(f"in-process run failed: {e!r}\n")
        code = 1
    ms = int((time.time() - t0) * 1000)
    gflops, rel = parse_metrics(buf.getvalue())
    return {"ms": ms, "exit_code": code, "gflops": gflops, "rel_error": rel}


def run_one(sched: str, ngpu: int, N: int, NB: int, seed: int = 42, timeout: float | None = None) -> dict:
    from .grid import ProcessGrid
    g = ProcessGrid.for_world(ngpu)
    env = dict(os.environ)
    env["STARPU_SCHED"] = sched
    env["CHOL_LOOKAHEAD"] = "0" if sched == "inorder" else "1"
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):  # benchmark.c:173-175
        env[k] = "1"
    args = driver_argv(0, ngpu, N, NB, g.P, g.Q, seed)
    if ngpu > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ngpu}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000),
               "-m", "dense_linear_app_b200.v6_test"] + args
    else:
        cmd = [sys.executable, "-m", "dense_linear_app_b200.v6_test"] + args
    t0 = time.time()
    try:
        pr = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, text=True, timeout=timeout)
        out, code = pr.stdout[:65536], pr.returncode
        if code < 0:
            code = 128 - code
    except subprocess.TimeoutExpired as e:
        out, code = (e.stdout or b"").decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or ""), 124
    ms = int((time.time() - t0) * 1000)
    gflops, rel = parse_metrics(out)
    return {"ms": ms, "exit_code": code, "gflops": gflops, "rel_error": rel}


def main(argv: list[str] | None = None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--N", type=int, nargs="*", default=list(NS))
    ap.add_argument("--NB", type=int, nargs="*", default=list(NBS))
    ap.add_argument("--ngpu", type=int, nargs="*", default=[1])
    ap.add_argument("--sched", nargs="*", default=list(SCHEDS))
    ap.add_argument("--repeats", type=int, default=REPEATS)
    ap.add_argument("--csv", default="results/bench.csv")
    ap.add_argument("--in-process", action="store_true",
                    help="1 GPU only: call the driver in this process instead of forking one child per run")
    a = ap.parse_args(argv)
    os.makedirs(os.path.dirname(a.csv) or ".", exist_ok=True)
    total = len(a.N) * len(a.NB) * len(a.ngpu) * len(a.sched) * a.repeats
    step = 0
    with open(a.csv, "a") as csv:
        if csv.tell() == 0:
            csv.write(HEADER)
            csv.flush()
        for N in a.N:
            for NB in a.NB:
                for ngpu in a.ngpu:
                    mapping = f"{ngpu}_b200"
                    for sched in a.sched:
                        for r in range(a.repeats):
                            step += 1
                            print(f"------------------ sched={sched} N={N} NB={NB} mapping={mapping} étape {step} sur "
                                  f"{total} -----------------------------", flush=True)
                            res = run_in_process(sched, N, NB) if (a.in_process and ngpu == 1) else run_one(sched, ngpu, N, NB)
                            ts = time.strftime("%Y-%m-%d %H:%M:%S")
                            csv.write(f"{ts},{sched},{mapping},0,{ngpu},{N},{NB},{r},{res['ms']},{res['exit_code']},"
                                      f"{res['gflops']:.6f},{res['rel_error']:.6e}\n")
                            csv.flush()
                            print(f"   -> ms={res['ms']}  GF={res['gflops']:.2f}  err={res['rel_error']:.2e}  "
                                  f"exit={res['exit_code']}", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
