"""2D block-cyclic ownership of the lower tiles of an SPD matrix over a P x Q grid of GPUs.

The reference carries the grid in its descriptor (``p``, ``q`` arguments of
``CHAMELEON_Desc_Create``, v6_test.c:44-45) but only ever runs 1 x 1 (benchmark.c:130); here
``p x q`` is the grid of B200s of one box.  Pure host logic (no CUDA), so it is covered by the
CPU tests.

Tile (i, j), i >= j, lives on rank ``(i % P) * Q + (j % Q)``.  Each rank stores its tiles
column-packed: local columns in ascending j, inside a column ascending i, so the tiles a rank
owns in panel column k are contiguous (one NCCL broadcast per owner, no packing).
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache

import numpy as np


@dataclass(frozen=True)
class ProcessGrid:
    P: int = 1
    Q: int = 1

    def __post_init__(self):
        if self.P < 1 or self.Q < 1:
            raise ValueError("grid dimensions must be positive")

    @property
    def size(self) -> int:
        return self.P * self.Q

    def coords(self, rank: int) -> tuple[int, int]:
        return rank // self.Q, rank % self.Q

    def rank_of(self, p: int, q: int) -> int:
        return p * self.Q + q

    def owner(self, i: int, j: int) -> int:
        return (i % self.P) * self.Q + (j % self.Q)

    @staticmethod
    def for_world(world: int) -> "ProcessGrid":
        """Default grid for `world` GPUs: the most square P x Q with P <= Q (1x1, 1x2, 2x2, 2x4)."""
        p = int(np.floor(np.sqrt(world)))
        while world % p:
            p -= 1
        return ProcessGrid(p, world // p)


class LocalLayout:
    """Which lower tiles one rank owns and where they sit in its tile buffer."""

    def __init__(self, nt: int, grid: ProcessGrid, rank: int):
        self.nt, self.grid, self.rank = nt, grid, rank
        self.p, self.q = grid.coords(rank)
        P, Q = grid.P, grid.Q
        self.cols = [j for j in range(nt) if j % Q == self.q]
        self.col_start: dict[int, int] = {}   # j -> index of the first local tile of column j
        self.col_first_row: dict[int, int] = {}  # j -> smallest owned i >= j
        n = 0
        for j in self.cols:
            i0 = j + ((self.p - j) % P)
            self.col_start[j] = n
            self.col_first_row[j] = i0
            if i0 < nt:
                n += (nt - 1 - i0) // P + 1
        self.ntiles = n

    def owns(self, i: int, j: int) -> bool:
        return i % self.grid.P == self.p and j % self.grid.Q == self.q

    def index(self, i: int, j: int) -> int:
        """Local tile index of owned tile (i, j), i >= j."""
        if not (0 <= j <= i < self.nt) or not self.owns(i, j):
            raise KeyError((i, j))
        return self.col_start[j] + (i - self.col_first_row[j]) // self.grid.P

    def rows_in_col(self, j: int, above: int) -> range:
        """Owned row indices i > `above` (and >= j) of column j, ascending."""
        lo = max(j, above + 1)
        i0 = lo + ((self.p - lo) % self.grid.P)
        return range(i0, self.nt, self.grid.P)

    def tiles(self):
        """All owned (i, j) in storage order."""
        for j in self.cols:
            for i in range(self.col_first_row[j], self.nt, self.grid.P):
                yield i, j


@lru_cache(maxsize=None)
def panel_slots(nt: int, P: int, k: int) -> tuple[dict, list]:
    """Placement of panel column k (tiles (i, k), i > k) in a receive buffer grouped by owner row.

    Returns ({i: slot}, [(p, first_slot, count)] for p in 0..P-1): the tiles owned by process
    row p occupy `count` consecutive slots from `first_slot`, in ascending i — the same order
    they have in the owner's local storage.
    """
    slot, groups, n = {}, [], 0
    for p in range(P):
        lo = k + 1
        i0 = lo + ((p - lo) % P)
        rows = range(i0, nt, P)
        groups.append((p, n, len(rows)))
        for i in rows:
            slot[i] = n
            n += 1
    return slot, groups
