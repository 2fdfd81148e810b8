"""Process / device setup: the CHAMELEON_Init(ncpu, ngpu) / CHAMELEON_Finalize pair (v6_test.c:41,93;
worker_distrib.cpp:584-589) for one process per B200.

Under ``torch.distributed.run`` (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* in the environment) it
binds the process to its GPU and creates the NCCL process group used for the panel broadcasts,
with high-priority communication streams so the broadcast kernels are not queued behind the
trailing-update grid.
"""
from __future__ import annotations

import datetime
import os

import torch

from . import _lib

_state = {"inited": False, "rank": 0, "world": 1, "device": None}


def env_int(key: str, default: int) -> int:
    """env_int (worker_distrib.cpp:82-85): non-negative integer from the environment."""
    try:
        return max(0, int(os.environ[key]))
    except (KeyError, ValueError):
        return default


def init(ncpu: int = 0, ngpu: int | None = None) -> tuple[int, int]:
    """Returns (rank, world).  `ncpu` is accepted for signature parity and ignored (no CPU workers
    exist).  The number of GPUs is the number of launched ranks (WORLD_SIZE); `ngpu` — or env
    CHM_NGPU, the worker's knob (worker_distrib.cpp:585) — is only checked against it."""
    if _state["inited"]:
        return _state["rank"], _state["world"]
    if not torch.cuda.is_available():
        raise RuntimeError("dense-linear-app_b200 needs a CUDA device (B200): there is no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    want = ngpu if ngpu else env_int("CHM_NGPU", 0)
    if want and want != world and rank == 0:
        import sys
        sys.stderr.write(f"[setup] ngpu={want} requested but {world} rank(s) launched: using {world} GPU(s)\n")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _lib.call("chol_init", local)
    if world > 1:
        # NCCL writes its banner ("NCCL version ...", when NCCL_DEBUG is set) to stdout by default;
        # stdout carries the driver's parsed lines (v6_test.c:64,86) / bench.py's JSON line
        # (NCCL_DEBUG_FILE is ignored at level VERSION, so that level is dropped; other levels log to a file)
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            del os.environ["NCCL_DEBUG"]
        elif os.environ.get("NCCL_DEBUG"):
            os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug.%h.%p.log")
        # A receiving rank joins the panel broadcast early and its NCCL kernel then spins on one SM
        # per channel until the owner has factored the panel.  CHOL_NCCL_CHANNELS=n caps the
        # channels (measured on 8 B200: n=4 lifts the update kernel from 28.4 to 31.9 TFLOP/s but
        # slows the broadcasts so much that the whole run drops from 204 to 183 TFLOP/s), so the
        # default (0) leaves NCCL's own choice; an explicit NCCL_MAX_NCHANNELS wins.
        nch = os.environ.get("CHOL_NCCL_CHANNELS", "0")
        if nch != "0":
            os.environ.setdefault("NCCL_MAX_NCHANNELS", nch)
        import torch.distributed as dist
        if not dist.is_initialized():
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            dist.init_process_group("nccl", rank=rank, world_size=world, pg_options=opts,
                                    timeout=datetime.timedelta(seconds=int(os.environ.get("CHOL_NCCL_TIMEOUT_S", "300"))),
                                    device_id=torch.device("cuda", local))
    _state.update(inited=True, rank=rank, world=world, device=torch.device("cuda", local))
    return rank, world


def finalize() -> None:
    """CHAMELEON_Finalize (v6_test.c:93)."""
    if _state["inited"] and _state["world"] > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    _state["inited"] = False
