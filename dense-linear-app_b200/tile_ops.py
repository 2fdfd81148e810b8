"""The four tile operations of the ArmoniK worker as Python functions on device tiles.

Each function takes column-major b x b FP64 tiles held in CUDA tensors (see tiles.py for how a
column-major tile looks to torch), works in place and is asynchronous on the current stream —
the same contract as the Chameleon calls they replace on 1-tile descriptors:

    potrf_tile  CHAMELEON_dpotrf_Tile(ChamLower, dA)                                  worker_distrib.cpp:238
    trsm_tile   CHAMELEON_dtrsm_Tile(ChamRight, ChamLower, ChamTrans, ChamNonUnit, 1, dL, dA)    :323
    syrk_tile   CHAMELEON_dsyrk_Tile(ChamLower, ChamNoTrans, -1, dA, 1, dC)                      :416
    gemm_tile   CHAMELEON_dgemm_Tile(ChamNoTrans, ChamTrans, -1, dAi, dAj, 1, dC)                :511
"""
from __future__ import annotations

import torch

from . import _lib

_workspaces: dict = {}


def _check(t: torch.Tensor, name: str) -> int:
    if not t.is_cuda:
        raise _lib.CholError(name, -1, "tile must live on a CUDA device: this library has no CPU path")
    if t.dtype != torch.float64 or t.dim() != 2 or t.shape[0] != t.shape[1] or not t.is_contiguous():
        raise ValueError(f"{name}: tile must be a contiguous square float64 tensor")
    return t.shape[0]


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _workspace(b: int, device: torch.device) -> torch.Tensor:
    """Per-device scratch for the inverted diagonal blocks (grown on demand, reused; stream
    ordered on the caller's stream like the tile ops themselves)."""
    key = (device.index, b)
    w = _workspaces.get(key)
    if w is None:
        nbytes = max(_lib.load().chol_potrf_tile_workspace(b), 8)
        w = torch.empty(nbytes // 8, dtype=torch.float64, device=device)
        _workspaces[key] = w
    return w


def potrf_tile(A: torch.Tensor, info: torch.Tensor | None = None) -> torch.Tensor:
    """A <- chol_lower(A); the strict upper triangle is left untouched.  Returns the device int32
    tensor holding LAPACK info (0, or 1-based index of the first non-positive pivot)."""
    b = _check(A, "potrf_tile")
    if info is None:
        info = torch.zeros(1, dtype=torch.int32, device=A.device)
    _lib.call("chol_potrf_tile", b, A.data_ptr(), b, _workspace(b, A.device).data_ptr(), info.data_ptr(), 0,
              _stream(A))
    return info


def trsm_tile(L: torch.Tensor, A: torch.Tensor) -> None:
    """A <- A L^{-T} (Right, Lower, Trans, NonUnit, alpha = 1); only the lower triangle of L is read."""
    b = _check(A, "trsm_tile")
    if _check(L, "trsm_tile") != b:
        raise ValueError("trsm_tile: L and A must have the same size")
    _lib.call("chol_trsm_tile", b, L.data_ptr(), b, A.data_ptr(), b, _workspace(b, A.device).data_ptr(), _stream(A))


def syrk_tile(A: torch.Tensor, C: torch.Tensor) -> None:
    """C <- C - A A^T, lower triangle of C only."""
    b = _check(C, "syrk_tile")
    if _check(A, "syrk_tile") != b:
        raise ValueError("syrk_tile: A and C must have the same size")
    _lib.call("chol_syrk_tile", b, A.data_ptr(), b, C.data_ptr(), b, _stream(C))


def gemm_tile(Ai: torch.Tensor, Aj: torch.Tensor, C: torch.Tensor) -> None:
    """C <- C - Ai Aj^T."""
    b = _check(C, "gemm_tile")
    if _check(Ai, "gemm_tile") != b or _check(Aj, "gemm_tile") != b:
        raise ValueError("gemm_tile: Ai, Aj and C must have the same size")
    _lib.call("chol_gemm_tile", b, Ai.data_ptr(), b, Aj.data_ptr(), b, C.data_ptr(), b, _stream(C))


def potrf_batched(A: torch.Tensor) -> torch.Tensor:
    """`batch` independent lower Cholesky factorizations of the n x n column-major matrices
    A[batch, n, n] (the many-small-tasks workload: one POTRF task per tile,
    client_distrib.cpp v1:139-141).  Returns the per-matrix info tensor (int32, device)."""
    if not A.is_cuda:
        raise _lib.CholError("potrf_batched", -1, "matrices must live on a CUDA device: no CPU path")
    if A.dtype != torch.float64 or A.dim() != 3 or A.shape[1] != A.shape[2] or not A.is_contiguous():
        raise ValueError("potrf_batched: A must be a contiguous float64 tensor [batch, n, n]")
    batch, n = A.shape[0], A.shape[1]
    info = torch.zeros(max(batch, 1), dtype=torch.int32, device=A.device)
    _lib.call("chol_potrf_batched", n, batch, A.data_ptr(), n, n * n, info.data_ptr(), _stream(A))
    return info[:batch]


def potrf_batched_from_host(host_in: torch.Tensor, host_out: torch.Tensor, host_info: torch.Tensor,
                            device: torch.device | None = None, chunks: int = 16) -> None:
    """End-to-end form of potrf_batched for matrices that live in (pinned) HOST memory — the many-task
    situation of the ArmoniK worker, where every tile arrives as a host blob (worker_distrib.cpp:186,261).
    The batch is cut into `chunks` pieces that flow through three streams: upload of piece i+1, factorization
    of piece i and download of piece i-1 overlap, so the step costs about max(PCIe in, PCIe out, kernel)
    instead of their sum.  Asynchronous on the current stream; synchronise before reading host_out."""
    if not (host_in.is_pinned() and host_out.is_pinned() and host_info.is_pinned()):
        raise ValueError("potrf_batched_from_host: host tensors must be pinned")
    if host_in.dtype != torch.float64 or host_in.dim() != 3 or host_in.shape[1] != host_in.shape[2]:
        raise ValueError("potrf_batched_from_host: host_in must be float64 [batch, n, n]")
    if host_out.shape != host_in.shape or host_out.dtype != torch.float64:
        raise ValueError("potrf_batched_from_host: host_out must have the shape and dtype of host_in")
    if host_info.dtype != torch.int32 or host_info.numel() < host_in.shape[0]:
        raise ValueError("potrf_batched_from_host: host_info must be int32 with one entry per matrix")
    dev = device or torch.device("cuda", torch.cuda.current_device())
    batch, n = host_in.shape[0], host_in.shape[1]
    if batch == 0 or n == 0:
        return
    per = -(-batch // max(1, chunks))
    key = (dev.index, per, n)
    st = _pipelines.get(key)
    if st is None:
        st = _pipelines[key] = {"buf": [torch.empty((per, n, n), dtype=torch.float64, device=dev) for _ in range(3)],
                                "info": [torch.zeros(per, dtype=torch.int32, device=dev) for _ in range(3)],
                                "up": torch.cuda.Stream(dev), "run": torch.cuda.Stream(dev), "down": torch.cuda.Stream(dev),
                                "free": [None, None, None]}
    cur = torch.cuda.current_stream(dev)
    for s_ in (st["up"], st["run"], st["down"]):
        s_.wait_stream(cur)
    for i, lo in enumerate(range(0, batch, per)):
        hi = min(batch, lo + per)
        slot = i % 3
        buf, inf = st["buf"][slot][:hi - lo], st["info"][slot][:hi - lo]
        with torch.cuda.stream(st["up"]):
            if st["free"][slot] is not None:
                st["up"].wait_event(st["free"][slot])          # the download that last used this slot
            buf.copy_(host_in[lo:hi], non_blocking=True)
            ev_up = torch.cuda.Event()
            ev_up.record(st["up"])
        with torch.cuda.stream(st["run"]):
            st["run"].wait_event(ev_up)
            _lib.call("chol_potrf_batched", n, hi - lo, buf.data_ptr(), n, n * n, inf.data_ptr(), st["run"].cuda_stream)
            ev_run = torch.cuda.Event()
            ev_run.record(st["run"])
        with torch.cuda.stream(st["down"]):
            st["down"].wait_event(ev_run)
            host_out[lo:hi].copy_(buf, non_blocking=True)
            host_info[lo:hi].copy_(inf, non_blocking=True)
            st["free"][slot] = torch.cuda.Event()
            st["free"][slot].record(st["down"])
    for s_ in (st["up"], st["run"], st["down"]):
        cur.wait_stream(s_)


_pipelines: dict = {}
