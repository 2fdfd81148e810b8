"""``python -m dense_linear_app_b200.v3_cli --N .. --NB .. ...`` — the named-argument driver.

Mirrors the reference's ``v3_script_cholesky_x_arg_gpt.c`` (Cholesky_Chameleon_sauv/code_c; V3 below):
the 20 long options, all required (V3:69-92,131-135), the strict geometry checks with the same
messages (V3:178-199), stdout ``N=.. NB=.. ...`` / ``Time: %.6f s`` / ``Performance: %.2f Gflop/s``
(V3:237-240) and exit status ``info != 0`` (V3:247).  ``--bump`` and ``--seed`` feed the dplgsy-style
generator.  Only ``--dtyp d`` and ``--uplo L`` run: the reference maps s/z/c and U/B but calls the
``d`` routines for every type (V3:226), and nothing in it ever uses U.
"""
from __future__ import annotations

import sys
import time

OPTIONS = ("N", "NB", "ncpu", "ngpu", "mat", "dtyp", "mb", "nb", "bsiz", "lm", "ln", "i", "j", "m", "n", "p", "q",
           "bump", "uplo", "seed")
CHAM_UPLO = {"U": 121, "L": 122, "B": 123}     # ChamUpper / ChamLower / ChamUpperLower


def usage(prog: str) -> str:
    return (f"Usage: {prog} --N INT --NB INT --ncpu INT --ngpu INT --mat none|user --dtyp d|s|z|c \\\n"
            "          --mb INT --nb INT --bsiz INT --lm INT --ln INT --i INT --j INT \\\n"
            "          --m INT --n INT --p INT --q INT --bump DOUBLE --uplo L|U|B --seed ULL\n\n"
            "ALL options are required. No defaults.\n\n"
            f"Example:\n  {prog} --N 3000 --NB 256 --ncpu 4 --ngpu 1 --mat none --dtyp d \\\n"
            "     --mb 256 --nb 256 --bsiz 65536 --lm 3000 --ln 3000 --i 0 --j 0 \\\n"
            "     --m 3000 --n 3000 --p 1 --q 1 --bump 3000 --uplo L --seed 51\n")


def _strtol(s: str) -> int:
    """strtol(s, NULL, 10): leading integer prefix, 0 if none (V3:48)."""
    s = s.strip()
    for n in range(len(s), 0, -1):
        try:
            return int(s[:n])
        except ValueError:
            continue
    return 0


def map_dtyp(s: str):
    """map_dtyp_from_string (V3:25-34) -> 'd' | 's' | 'z' | 'c' | None."""
    return {"d": "d", "D": "d", "0": "d", "s": "s", "S": "s", "1": "s", "z": "z", "Z": "z", "2": "z", "c": "c",
            "C": "c", "3": "c"}.get(s)


def map_uplo(s: str):
    """map_uplo_from_string (V3:36-44) -> 'L' | 'U' | 'B' | None."""
    return {"L": "L", "l": "L", "0": "L", "U": "U", "u": "U", "1": "U", "B": "B", "b": "B", "2": "B"}.get(s)


def parse(argv: list[str], err=sys.stderr):
    """getopt_long loop + the checks of V3:94-199.  Returns (exit_code, None) or (None, dict)."""
    prog, vals, it = argv[0], {}, iter(argv[1:])
    for arg in it:
        if arg in ("-h", "--help"):
            err.write(usage(prog))
            return 0, None
        if not arg.startswith("--"):
            continue                                   # getopt_long skips non-option words
        name, eq, val = arg[2:].partition("=")
        if name not in OPTIONS:
            err.write(usage(prog))
            return 1, None
        if not eq:
            val = next(it, None)
            if val is None:
                err.write(usage(prog))
                return 1, None
        vals[name] = val
    if any(o not in vals for o in OPTIONS):
        err.write("Error: all options are required. Missing at least one.\n" + usage(prog))
        return 1, None
    a = {k: _strtol(vals[k]) for k in ("N", "NB", "ncpu", "ngpu", "mb", "nb", "bsiz", "lm", "ln", "i", "j", "m", "n",
                                       "p", "q")}
    try:
        a["bump"] = float(vals["bump"])
    except ValueError:
        a["bump"] = 0.0
    a["seed"] = _strtol(vals["seed"]) & 0xFFFFFFFFFFFFFFFF
    a["dtyp"], a["uplo"] = map_dtyp(vals["dtyp"]), map_uplo(vals["uplo"])
    a["mat_user"] = vals["mat"] not in ("none", "NULL", "0")
    if a["dtyp"] is None:
        err.write(f"Error: invalid --dtyp {vals['dtyp']}\n")
        return 1, None
    if a["uplo"] is None:
        err.write(f"Error: invalid --uplo {vals['uplo']}\n")
        return 1, None
    if min(a[k] for k in ("N", "NB", "mb", "nb", "lm", "ln", "m", "n", "p", "q")) <= 0:
        err.write("Error: dimension arguments must be >0.\n")
        return 1, None
    if a["bsiz"] < a["mb"] * a["nb"]:
        err.write(f"Error: --bsiz < mb*nb (bsiz={a['bsiz']} mb={a['mb']} nb={a['nb']}).\n")
        return 1, None
    if a["i"] < 0 or a["j"] < 0 or a["i"] >= a["lm"] or a["j"] >= a["ln"]:
        err.write(f"Error: invalid offsets i={a['i']} j={a['j']} (lm={a['lm']} ln={a['ln']}).\n")
        return 1, None
    if a["i"] + a["m"] > a["lm"] or a["j"] + a["n"] > a["ln"]:
        err.write(f"Error: submatrix (i={a['i']},m={a['m']}) outside lm={a['lm']} OR (j={a['j']},n={a['n']}) outside "
                  f"ln={a['ln']}.\n")
        return 1, None
    if a["bump"] == 0.0:
        err.write("Warning: bump==0 -> matrix may not be SPD.\n")
    return None, a


def main(argv: list[str] | None = None) -> int:
    argv = sys.argv if argv is None else argv
    code, a = parse(argv)
    if code is not None:
        return code
    if a["dtyp"] != "d":
        sys.stderr.write("Error: only --dtyp d runs (the reference calls the d routines for every type).\n")
        return 1
    if a["uplo"] not in ("L", "U"):
        # ChamUpperLower makes no sense for a Cholesky factorization (Chameleon returns -1 for it)
        sys.stderr.write("Error: --uplo must be L (ChamLower) or U (ChamUpper) for dpotrf.\n")
        return 1

    import torch
    from . import runtime
    from .cholesky import TiledCholesky, transpose_tiles
    from .tiles import TileDesc, TileMatrix

    rank, world = runtime.init(a["ncpu"], a["ngpu"])
    if a["p"] * a["q"] != world:
        sys.stderr.write(f"p*q = {a['p'] * a['q']} but {world} rank(s) were launched\n")
        return 1
    desc = TileDesc(a["mb"], a["nb"], a["bsiz"], a["lm"], a["ln"], a["i"], a["j"], a["m"], a["n"], a["p"], a["q"])
    try:
        A = TileMatrix(desc, rank).generate(a["bump"], a["seed"])
    except ValueError as e:
        sys.stderr.write(f"Error: {e}\n")
        return 1
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    # task submission (the plan) is part of CHAMELEON_dpotrf_Tile's bracket (v3_script_cholesky_x_arg_gpt.c:224-228)
    upper = a["uplo"] == "U"
    if upper:
        transpose_tiles(A)          # dplgsy(ChamUpper): storage position (i, j) now holds the upper tile (j, i)
        torch.cuda.synchronize()
    t0 = time.monotonic()
    if upper:
        transpose_tiles(A)          # A = U^T U as the lower factorization of the transposed tiles (cholesky.py)
    ch = TiledCholesky(A)
    ch.factor()
    if upper:
        transpose_tiles(A)
    info = ch.info()
    time_sec = time.monotonic() - t0
    dim = float(min(a["m"], a["n"]))
    gflops = (1.0 / 3.0) * dim ** 3 / (time_sec * 1e9)
    if rank == 0:
        print(f"N={a['N']} NB={a['NB']} ncpu={a['ncpu']} ngpu={a['ngpu']} p={a['p']} q={a['q']} bump={a['bump']:g} "
              f"uplo={CHAM_UPLO[a['uplo']]} seed={a['seed']}")
        print(f"Time: {time_sec:.6f} s")
        print(f"Performance: {gflops:.2f} Gflop/s", flush=True)
    if info != 0:
        sys.stderr.write(f"Erreur dans CHAMELEON_dpotrf_Tile: {info}\n")
    runtime.finalize()
    return int(info != 0)


if __name__ == "__main__":
    sys.exit(main())
