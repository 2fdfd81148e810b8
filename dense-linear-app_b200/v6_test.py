"""``python -m dense_linear_app_b200.v6_test <16 ints>`` — the whole-matrix driver.

Same command line, stdout lines and exit status as the reference's ``v6_test`` (v6_test.c):
    argv   ncpu ngpu N NB mb nb bsiz lm ln ioff joff m n p q seed          (v6_test.c:8-28)
    stdout "[setup] ...", "N = %d, NB = %d", "Time: %.3f s", "Performance: %.2f Gflop/s",
           "||A - LL^T||_inf / ||A||_inf = %.2e", "Validation numérique : ..."   (:35,62-64,86-87)
    exit   info != 0                                                               (:95)
so the sweep harness (bench_sweep.py, benchmark.c:45-67) parses it unchanged.  The residual is
the true ||A - L L^T||_inf / ||A||_inf (the reference's dlauum forms L^T L).  With
p*q > 1 launch it under torch.distributed.run with p*q ranks (one per GPU); rank 0 prints.
"""
from __future__ import annotations

import os
import sys
import time


USAGE = ("Usage: %s <num_cpu> <num_gpu> <matrix_size_N> <tile_size_NB> <mb> <nb> <bsiz> <lm> <ln> <ioff> <joff> "
         "<m> <n> <p> <q> <seed>\n")


def _atoi(s: str) -> int:
    """C atoi: leading integer prefix, 0 if none."""
    s = s.strip()
    n = 0
    for n in range(len(s), 0, -1):
        try:
            return int(s[:n])
        except ValueError:
            continue
    return 0


def main(argv: list[str] | None = None) -> int:
    argv = sys.argv if argv is None else argv
    if len(argv) < 17:
        sys.stderr.write(USAGE % argv[0])
        return 1
    (ncpu, ngpu, N, NB, mb, nb, bsiz, lm, ln, ioff, joff, m, n, p, q, seed) = [_atoi(a) for a in argv[1:17]]
    if bsiz != mb * nb:
        sys.stderr.write(f"Warning: bsiz ({bsiz}) != mb*nb ({mb * nb})\n")

    import torch
    from . import runtime
    from .cholesky import TiledCholesky
    from .tiles import TileDesc, TileMatrix

    rank, world = runtime.init(ncpu, ngpu)
    out = (lambda s: print(s, flush=True)) if rank == 0 else (lambda s: None)
    sched = os.environ.get("STARPU_SCHED")
    out(f"[setup] ncpu={ncpu} ngpu={ngpu} N={N} NB={NB} scheduler={sched if sched else '(default)'}")
    if p * q != world:
        sys.stderr.write(f"p*q = {p * q} but {world} rank(s) were launched\n")
        return 1
    desc = TileDesc(mb, nb, max(bsiz, mb * nb), lm, ln, ioff, joff, m, n, p, q)
    A = TileMatrix(desc, rank).generate(float(N), seed)                      # dplgsy(bump = N)
    Aorig = A.clone()                                                        # dlacpy(UpperLower)
    lookahead = os.environ.get("CHOL_LOOKAHEAD", "1") != "0"
    if world == 1:
        # what CHAMELEON_Init does for its codelets (cuBLAS handles, v6_test.c:41): load the kernels of this
        # tile size once, outside the timed region, on a 2 x 2-tile dummy
        W = TileMatrix(TileDesc.square(2 * mb, mb), 0).generate(float(2 * mb), 1)
        TiledCholesky(W, lookahead=lookahead).factor()
        del W
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    # the reference times all of CHAMELEON_dpotrf_Tile, task submission included (v6_test.c:54-57):
    # building the plan (the DAG as device task lists) is inside the bracket
    t0 = time.monotonic()
    ch = TiledCholesky(A, lookahead=lookahead)
    t_plan = time.monotonic() - t0
    ch.factor()
    info = ch.info()                                                          # synchronises
    t1 = time.monotonic()
    time_sec = t1 - t0
    gflops = (1.0 / 3.0) * float(N) ** 3 / (time_sec * 1e9)
    out(f"N = {N}, NB = {NB}")
    out(f"Time: {time_sec:.3f} s")
    out(f"[plan] {t_plan * 1e3:.1f} ms of it building and uploading the task lists")
    out(f"Performance: {gflops:.2f} Gflop/s")
    if info != 0:
        sys.stderr.write(f"Erreur dans CHAMELEON_dpotrf_Tile: {info}\n")
    rel = ch.residual(Aorig)["inf"] if info == 0 else float("nan")
    out(f"||A - LL^T||_inf / ||A||_inf = {rel:.2e}")
    out("Validation numérique : %s" % ("✅ PASS" if rel < 1e-10 else "❌ FAIL"))
    runtime.finalize()
    return int(info != 0)


if __name__ == "__main__":
    sys.exit(main())
