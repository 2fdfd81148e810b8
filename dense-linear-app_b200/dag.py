"""Client side of the tile-task interface: parameters, SPD generator, tile extraction, JSON
payloads and the right-looking wave loop, mirroring the ArmoniK client
(``client_distrib.cpp``; C1 = w_c_cons_v1/client_construction, C2 = w_c_cons_v2/client_construction2).

Pure host code (numpy), no CUDA: this is the DAG the Python side builds.  Op routing follows v1
(explicit op names, C1:44-97); v2's id-prefix heuristic (C2:165-194) turned every TRSM into a
SYRK and is not reproduced.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Callable, Iterator, Mapping, Sequence

import numpy as np


# ---- parameters (C2:58-93) ---------------------------------------------------------------------
@dataclass
class Params:
    N: int = 12
    B: int = 4


def _parse_int(s: str, fallback: int, what: str) -> int:
    try:
        return int(s)
    except ValueError:
        print(f"[WARN] invalid integer for {what}: '{s}' -> using fallback {fallback}")
        return fallback


def load_params(argv: Sequence[str] = (), env: Mapping[str, str] | None = None) -> Params:
    """Env CHOLESKY_N / CHOLESKY_B first, then ``--N=.. --B=..`` flags or two positionals N B."""
    env = os.environ if env is None else env
    p = Params()
    if "CHOLESKY_N" in env:
        p.N = _parse_int(env["CHOLESKY_N"], p.N, "CHOLESKY_N")
    if "CHOLESKY_B" in env:
        p.B = _parse_int(env["CHOLESKY_B"], p.B, "CHOLESKY_B")
    seen = 0
    for arg in argv:
        if arg.startswith("--N="):
            p.N = _parse_int(arg[4:], p.N, "--N")
        elif arg.startswith("--B="):
            p.B = _parse_int(arg[4:], p.B, "--B")
        elif arg and not arg.startswith("-"):
            if seen == 0:
                p.N = _parse_int(arg, p.N, "N")
            elif seen == 1:
                p.B = _parse_int(arg, p.B, "B")
            seen += 1
    if p.N <= 0 or p.B <= 0:
        raise ValueError("N and B must be positive")
    return p


# ---- std::mt19937_64 + uniform_real_distribution<double>(-0.5, 0.5) (C2:231-232) ----------------
class MT19937_64:
    """The 64-bit Mersenne Twister of <random> (numpy only ships the 32-bit one)."""
    NN, MM = 312, 156
    _A = np.uint64(0xB5026F5AA96619E9)
    _UM = np.uint64(0xFFFFFFFF80000000)
    _LM = np.uint64(0x7FFFFFFF)

    def __init__(self, seed: int):
        mt = np.empty(self.NN, dtype=np.uint64)
        x = seed & 0xFFFFFFFFFFFFFFFF
        mt[0] = x
        for i in range(1, self.NN):
            x = (6364136223846793005 * (x ^ (x >> 62)) + i) & 0xFFFFFFFFFFFFFFFF
            mt[i] = x
        self.mt, self.pos = mt, self.NN

    def _twist(self) -> None:
        mt, NN, MM = self.mt, self.NN, self.MM
        one = np.uint64(1)

        def mix(up, lo, far):
            x = (up & self._UM) | (lo & self._LM)
            return far ^ (x >> one) ^ np.where((x & one).astype(bool), self._A, np.uint64(0))

        mt[:NN - MM] = mix(mt[:NN - MM], mt[1:NN - MM + 1], mt[MM:NN])
        mt[NN - MM:NN - 1] = mix(mt[NN - MM:NN - 1], mt[NN - MM + 1:NN], mt[:MM - 1])
        mt[NN - 1:] = mix(mt[NN - 1:], mt[:1], mt[MM - 1:MM])
        self.pos = 0

    def raw(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint64)
        done = 0
        while done < n:
            if self.pos == self.NN:
                self._twist()
            take = min(n - done, self.NN - self.pos)
            out[done:done + take] = self.mt[self.pos:self.pos + take]
            self.pos += take
            done += take
        x = out
        x ^= (x >> np.uint64(29)) & np.uint64(0x5555555555555555)
        x ^= (x << np.uint64(17)) & np.uint64(0x71D67FFFEDA60000)
        x ^= (x << np.uint64(37)) & np.uint64(0xFFF7EEE000000000)
        x ^= x >> np.uint64(43)
        return x

    def uniform(self, n: int, a: float = -0.5, b: float = 0.5) -> np.ndarray:
        """n draws of std::uniform_real_distribution<double>(a, b) (libstdc++: one 64-bit word per
        draw, canonical = double(word) / 2^64, clamped below 1)."""
        c = self.raw(n).astype(np.float64) / 18446744073709551616.0
        c[c >= 1.0] = np.nextafter(1.0, 0.0)
        return (b - a) * c + a


class NormalDist:
    """std::normal_distribution<double>(0, 1) of libstdc++ on an MT19937_64: Marsaglia's polar method
    over generate_canonical<double, 53> (one 64-bit word per canonical draw), the second value of
    each accepted pair saved for the next call.  math.log / math.sqrt are libm's, as in the reference
    binary, so the stream reproduces the C++ one bit for bit (tests/golden/normal_dist.json)."""

    def __init__(self, gen: MT19937_64):
        self.gen, self.saved = gen, None

    def _canonical(self) -> float:
        c = float(self.gen.raw(1)[0]) / 18446744073709551616.0
        return c if c < 1.0 else float(np.nextafter(1.0, 0.0))

    def __call__(self) -> float:
        import math
        if self.saved is not None:
            v, self.saved = self.saved, None
            return v
        while True:
            x = 2.0 * self._canonical() - 1.0
            y = 2.0 * self._canonical() - 1.0
            r2 = x * x + y * y
            if not (r2 > 1.0 or r2 == 0.0):
                break
        mult = math.sqrt(-2.0 * math.log(r2) / r2)
        self.saved = x * mult
        return y * mult


_v1_normal: NormalDist | None = None


def generate_random_B_block(B: int, scale: float = 1.0, dist: NormalDist | None = None) -> np.ndarray:
    """generate_random_B_block (C1:102-108): B*B draws of N(0, 1) * scale from ONE process-wide
    mt19937_64(42) (the reference's generator is a function-level static, so consecutive calls
    continue the same stream; pass `dist` to use a private stream).  Returned as the flat tile
    blob the client uploads (B*B doubles; the worker reads it column-major, W1:212-227)."""
    global _v1_normal
    if dist is None:
        if _v1_normal is None:
            _v1_normal = NormalDist(MT19937_64(42))
        dist = _v1_normal
    return np.array([dist() * scale for _ in range(B * B)], dtype=np.float64)


def make_blocks_v1(N: int, B: int, dist: NormalDist | None = None) -> dict:
    """The v1 client's input (C1:189-192): every lower tile (i, j), i-major then j, is an independent
    N(0, 0.1^2) block, diagonal tiles get +B on their diagonal.  Diagonal tiles are therefore NOT
    symmetric: the factorization must read their lower triangle only (POTRF/SYRK uplo=Lower).
    Returns {(i, j): flat blob}."""
    dist = dist or NormalDist(MT19937_64(42))
    nb = N // B
    out = {}
    for i in range(nb):
        for j in range(i + 1):
            blk = generate_random_B_block(B, 0.1, dist)
            if i == j:
                blk[np.arange(B) * B + np.arange(B)] += float(B)
            out[(i, j)] = blk
    return out


def make_spd_like_chameleon(N: int, bump: float = 100.0, uplo: str = "L", seed: int = 12345) -> np.ndarray:
    """make_spd_like_chameleon (C2:224-252): lower triangle filled column by column with
    U(-0.5, 0.5) draws of mt19937_64(seed), mirrored, diagonal += bump.  Column-major N x N."""
    gen = MT19937_64(seed)
    A = np.zeros((N, N), order="F")
    vals = gen.uniform(N * (N + 1) // 2)
    if uplo in ("L", "l"):
        jj, ii = np.triu_indices(N)      # (j, i) pairs with i >= j, j-major then i ascending
        A[ii, jj] = vals
        A[jj, ii] = vals
    else:
        jj, ii = np.tril_indices(N)      # column j, rows i <= j: j-major then i ascending
        A[ii, jj] = vals
        A[jj, ii] = vals
    A[np.arange(N), np.arange(N)] += bump
    return A


def enforce_strict_diag_dominance(A: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    """enforce_strict_diag_dominance (C2:255-264), in place."""
    off = np.abs(A).sum(axis=1) - np.abs(np.diag(A))
    need = off + eps - np.diag(A)
    idx = np.arange(A.shape[0])
    A[idx, idx] += np.where(need > 0.0, need, 0.0)
    return A


def extract_block(A: np.ndarray, B: int, bi: int, bj: int) -> np.ndarray:
    """extract_block_from_spd_matrix_colmajor (C2:280-309): B x B column-major copy of block
    (bi, bj), zero-padded past the matrix edge."""
    N = A.shape[0]
    blk = np.zeros((B, B), order="F")
    r0, c0 = bi * B, bj * B
    r1, c1 = min(N, r0 + B), min(N, c0 + B)
    if r1 > r0 and c1 > c0:
        blk[: r1 - r0, : c1 - c0] = A[r0:r1, c0:c1]
    return blk


def block_id_from_ij(i: int, j: int) -> str:
    """Result name of tile (i, j) (C2:319-321)."""
    return f"blk/{i}/{j}"


# ---- payloads (C1:44-97) -------------------------------------------------------------------------
def _dumps(d: dict) -> str:
    return json.dumps(d, separators=(",", ":"))


def make_payload_potrf(id_in: str, B: int) -> str:
    return _dumps({"op": "POTRF", "B": B, "in": id_in})


def make_payload_trsm(id_Lkk: str, id_Aik: str, B: int) -> str:
    return _dumps({"op": "TRSM", "B": B, "inL": id_Lkk, "inA": id_Aik})


def make_payload_syrk(id_Cii: str, id_Aik: str, B: int) -> str:
    return _dumps({"op": "SYRK", "B": B, "inC": id_Cii, "inA": id_Aik})


def make_payload_gemm(id_Cij: str, id_Aik: str, id_Ajk: str, B: int) -> str:
    return _dumps({"op": "GEMM", "B": B, "inC": id_Cij, "inAi": id_Aik, "inAj": id_Ajk})


@dataclass(frozen=True)
class TileTask:
    """One node of the DAG: op on tile `out` = (i, j) reading tiles `deps` (first = the tile updated)."""
    op: str
    k: int
    out: tuple[int, int]
    deps: tuple[tuple[int, int], ...]


def build_dag(N: int, B: int) -> Iterator[TileTask]:
    """The wave loop (C1:278-333 / C2:506-565) as a task sequence, in the reference's submission
    order: for k: POTRF(k,k); TRSM(i,k) i>k; for i>k, k<j<=i: SYRK(i,i) if i==j else GEMM(i,j)."""
    nb = (N + B - 1) // B
    for k in range(nb):
        yield TileTask("POTRF", k, (k, k), ((k, k),))
        for i in range(k + 1, nb):
            yield TileTask("TRSM", k, (i, k), ((i, k), (k, k)))
        for i in range(k + 1, nb):
            for j in range(k + 1, i + 1):
                if i == j:
                    yield TileTask("SYRK", k, (i, i), ((i, i), (i, k)))
                else:
                    yield TileTask("GEMM", k, (i, j), ((i, j), (i, k), (j, k)))


def task_counts(N: int, B: int) -> dict:
    nb = (N + B - 1) // B
    return {"POTRF": nb, "TRSM": nb * (nb - 1) // 2, "SYRK": nb * (nb - 1) // 2,
            "GEMM": nb * (nb - 1) * (nb - 2) // 6}


def payload_for(task: TileTask, ids: Mapping[tuple[int, int], str], B: int) -> tuple[str, list[str]]:
    """(payload JSON, dependency ids) for a task given the current result id of every tile."""
    if task.op == "POTRF":
        a = ids[task.out]
        return make_payload_potrf(a, B), [a]
    if task.op == "TRSM":
        (i, k), kk = task.deps
        return make_payload_trsm(ids[kk], ids[(i, k)], B), [ids[kk], ids[(i, k)]]
    if task.op == "SYRK":
        c, a = task.deps
        return make_payload_syrk(ids[c], ids[a], B), [ids[c], ids[a]]
    c, ai, aj = task.deps
    return make_payload_gemm(ids[c], ids[ai], ids[aj], B), [ids[c], ids[ai], ids[aj]]


def run_waves(N: int, B: int, blocks: dict, submit_one: Callable[[str, dict], bytes]) -> dict:
    """The client main loop (C2:442-565) against any task executor: `blocks` maps "blk/i/j" to the
    tile blob (raw B*B little-endian doubles, column-major); every task is submitted and awaited
    one at a time (C2:498-499) through ``submit_one(payload_json, {dep id: blob}) -> blob``.
    Each result gets a fresh id, `latest` tracks the current one per tile.  Returns
    {"blk/i/j": final blob}."""
    store = dict(blocks)                                   # object store: id -> blob
    latest = {name: name for name in blocks}               # tile name -> current result id
    serial = 0
    for task in build_dag(N, B):
        ids = {t: latest[block_id_from_ij(*t)] for t in task.deps}
        payload, deps = payload_for(task, ids, B)
        out = submit_one(payload, {d: store[d] for d in deps})
        serial += 1
        out_id = f"out/{serial}"
        store[out_id] = out
        latest[block_id_from_ij(*task.out)] = out_id
    return {name: store[rid] for name, rid in latest.items()}
