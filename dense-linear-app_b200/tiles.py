"""Tile descriptor and device-resident tile storage.

``TileDesc`` mirrors the argument list of ``CHAMELEON_Desc_Create(&d, mat, ChamRealDouble, mb, nb,
bsiz, lm, ln, i, j, m, n, p, q)`` (v6_test.c:44-45; worker_distrib.cpp:76-79 for the 1-tile
form).  ``TileMatrix`` is the storage behind a descriptor: the lower tiles this rank owns, each
b x b column-major FP64 with ld = b (worker_distrib.cpp:212-227), in one torch buffer.

torch is used for the buffer only.  A column-major tile appears to torch as a (b, b) tensor
``t`` with ``t[c, r] = A(r, c)``, i.e. torch sees the transpose.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .grid import LocalLayout, ProcessGrid


@dataclass(frozen=True)
class TileDesc:
    """The 11 integers of a Chameleon descriptor (v6_test.c:17-27)."""
    mb: int
    nb: int
    bsiz: int
    lm: int
    ln: int
    i: int
    j: int
    m: int
    n: int
    p: int = 1
    q: int = 1

    def validate(self) -> None:
        """Geometry rules of the reference drivers (v3_script_cholesky_x_arg_gpt.c:178-196):
        positive sizes, bsiz >= mb*nb, the m x n sub-matrix at (i, j) inside lm x ln.  This
        framework additionally needs square tiles, a square matrix and zero offsets (the only
        geometry the reference ever runs, benchmark.c:123-130)."""
        for name in ("mb", "nb", "lm", "ln", "m", "n", "p", "q"):
            if getattr(self, name) <= 0:
                raise ValueError(f"descriptor field {name} must be positive")
        if self.i < 0 or self.j < 0:
            raise ValueError("descriptor offsets must be non-negative")
        if self.bsiz < self.mb * self.nb:
            raise ValueError(f"bsiz ({self.bsiz}) < mb*nb ({self.mb * self.nb})")
        if self.i + self.m > self.lm or self.j + self.n > self.ln:
            raise ValueError("sub-matrix (i, j, m, n) does not fit in lm x ln")
        if self.mb != self.nb:
            raise ValueError("Cholesky needs square tiles (mb == nb)")
        if self.m != self.n:
            raise ValueError("Cholesky needs a square matrix (m == n)")
        if self.i != 0 or self.j != 0:
            raise ValueError("descriptor offsets other than 0 are not supported")

    @staticmethod
    def square(N: int, NB: int, p: int = 1, q: int = 1) -> "TileDesc":
        """The descriptor benchmark.c:123-130 builds: mb=nb=NB, bsiz=NB*NB, lm=ln=m=n=N."""
        return TileDesc(NB, NB, NB * NB, N, N, 0, 0, N, N, p, q)

    @staticmethod
    def one_block(B: int) -> "TileDesc":
        """create_desc_1block (worker_distrib.cpp:76-79)."""
        return TileDesc(B, B, B * B, B, B, 0, 0, B, B, 1, 1)


class TileMatrix:
    """Lower tiles of an N x N symmetric matrix owned by one rank of a P x Q grid.

    Ragged edges (N % b != 0) are padded inside the last tile row/column with the identity
    (generator and from_numpy do it), so every tile op works on full b x b tiles and the padding
    factors to the identity.
    """

    def __init__(self, desc: TileDesc, rank: int = 0, device: torch.device | str | None = None):
        desc.validate()
        self.desc = desc
        self.N, self.b = desc.m, desc.mb
        self.nt = (self.N + self.b - 1) // self.b
        self.grid = ProcessGrid(desc.p, desc.q)
        if not 0 <= rank < self.grid.size:
            raise ValueError("rank outside the process grid")
        self.rank = rank
        self.layout = LocalLayout(self.nt, self.grid, rank)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.buf = torch.empty((max(self.layout.ntiles, 1), self.b, self.b), dtype=torch.float64, device=self.device)
        self.tile_bytes = self.b * self.b * 8

    # ---- addressing ------------------------------------------------------------------------
    def tile(self, i: int, j: int) -> torch.Tensor:
        return self.buf[self.layout.index(i, j)]

    def tile_ptr(self, i: int, j: int) -> int:
        return self.buf.data_ptr() + self.layout.index(i, j) * self.tile_bytes

    # ---- fill / read back ------------------------------------------------------------------
    def generate(self, bump: float, seed: int) -> "TileMatrix":
        """CHAMELEON_dplgsy_Tile(bump, ChamLower, desc, seed) (v6_test.c:46), tile by tile on the
        device; counter-based, so every rank fills its own tiles with no communication."""
        st = torch.cuda.current_stream(self.device).cuda_stream
        b = self.b
        for i, j in self.layout.tiles():
            _lib.call("chol_plgsy_tile", float(bump), b, b, self.tile_ptr(i, j), b, self.desc.lm, i * b, j * b,
                      self.N, seed & 0xFFFFFFFFFFFFFFFF, st)
        return self

    def from_numpy(self, A: np.ndarray) -> "TileMatrix":
        """Cut the lower triangle of the host matrix A (N x N) into this rank's tiles
        (extract_block_from_spd_matrix_colmajor, client_distrib.cpp:280-309; the padding of an
        edge tile is the identity instead of the reference's zeros, which made ragged diagonal
        tiles singular)."""
        N, b = self.N, self.b
        assert A.shape == (N, N)
        host = torch.empty((max(self.layout.ntiles, 1), b, b), dtype=torch.float64)
        hv = host.numpy()
        for i, j in self.layout.tiles():
            blk = np.zeros((b, b))
            r1, c1 = min(N, (i + 1) * b), min(N, (j + 1) * b)
            blk[: r1 - i * b, : c1 - j * b] = A[i * b:r1, j * b:c1]
            if i == j and r1 - i * b < b:
                pad = np.arange(r1 - i * b, b)
                blk[pad, pad] = 1.0
            hv[self.layout.index(i, j)] = blk.T  # torch view is the transpose of the col-major tile
        self.buf.copy_(host)
        return self

    def to_numpy(self) -> np.ndarray:
        """Owned tiles assembled into an N x N host array (other entries zero)."""
        N, b = self.N, self.b
        out = np.zeros((self.nt * b, self.nt * b))
        hv = self.buf.cpu().numpy()
        for i, j in self.layout.tiles():
            out[i * b:(i + 1) * b, j * b:(j + 1) * b] = hv[self.layout.index(i, j)].T
        return np.asfortranarray(out[:N, :N])

    def clone(self) -> "TileMatrix":
        other = TileMatrix(self.desc, self.rank, self.device)
        other.buf.copy_(self.buf)
        return other
