"""Tile-task executor with the interface of the ArmoniK worker ``DagCholeskyWorker::Execute``
(``worker_distrib.cpp``; W2 = w_c_cons_v2/worker_construction2/src, line numbers below are W2).

One task = one tile op: parse the JSON payload (W2:47-69), fetch the dependency blobs by id
(W2:180-186), check their size (W2:218-220), run POTRF / TRSM / SYRK / GEMM on the B200 through
``libchol_b200.so`` and hand back the updated tile as raw ``B*B`` little-endian doubles,
column-major, ld = B (W2:227,261).  Failures never raise out of ``Execute``: they come back as a
``ProcessStatus`` carrying the same message text as the reference (W2:194-195, 218-220, 243-244,
547-549, 558-563).

Host blobs are staged through pinned buffers; the arithmetic is CUDA only (no CPU path).
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Mapping

import numpy as np
import torch

from . import tile_ops


@dataclass
class ProcessStatus:
    """armonik::api::worker::ProcessStatus: ok, or an error with a message."""
    ok: bool = True
    details: str = ""

    @staticmethod
    def Ok() -> "ProcessStatus":
        return ProcessStatus(True, "")

    @staticmethod
    def Error(msg: str) -> "ProcessStatus":
        return ProcessStatus(False, msg)


@dataclass
class TaskHandler:
    """The slice of armonik::api::worker::TaskHandler the worker uses (W2:105,180-186,261)."""
    payload: str
    data_dependencies: Mapping[str, bytes]
    expected_results: list = field(default_factory=lambda: ["output"])
    results: dict = field(default_factory=dict)

    def getPayload(self) -> str:
        return self.payload

    def getDataDependencies(self) -> Mapping[str, bytes]:
        return self.data_dependencies

    def getExpectedResults(self) -> list:
        return self.expected_results

    def send_result(self, key: str, data: bytes) -> None:
        self.results[key] = data


@dataclass
class Parsed:
    """struct Parsed (W2:46)."""
    op: str = ""
    B: int = 0
    in_: str = ""
    inL: str = ""
    inA: str = ""
    inC: str = ""
    inAi: str = ""
    inAj: str = ""


def handle_json(payload: str) -> Parsed:
    """handle_json (W2:47-69).  A malformed payload raises (the reference's rapidjson asserts);
    Execute turns that into an "Exception: ..." status."""
    d = json.loads(payload)
    if not isinstance(d, dict):
        raise ValueError("payload is not a JSON object")
    p = Parsed(op=_get(d, "op", str), B=_get(d, "B", int))
    if p.op == "POTRF":
        p.in_ = _get(d, "in", str)
    elif p.op == "TRSM":
        p.inL, p.inA = _get(d, "inL", str), _get(d, "inA", str)
    elif p.op == "SYRK":
        p.inC, p.inA = _get(d, "inC", str), _get(d, "inA", str)
    elif p.op == "GEMM":
        p.inC, p.inAi, p.inAj = _get(d, "inC", str), _get(d, "inAi", str), _get(d, "inAj", str)
    return p


def _get(d: dict, key: str, typ):
    if key not in d:
        raise KeyError(f"payload field '{key}' missing")
    v = d[key]
    if typ is int and (isinstance(v, bool) or not isinstance(v, int)):
        raise TypeError(f"payload field '{key}' is not an integer")
    if typ is str and not isinstance(v, str):
        raise TypeError(f"payload field '{key}' is not a string")
    return v


# the reference's log tags differ per op and include its typos (W2:195,374,471)
_TAG = {"POTRF": "[Worker][POTF]", "TRSM": "[Worker][TRSM]", "SYRK": "[Worker][SYRK]", "GEMM": "[Worker][GEMM]"}
_MISSING_PREFIX = {("SYRK", 0): " [Worker][SYRK]Missing dependency: ", ("GEMM", 0): " [Worker][GEMM] Missing dependency: "}


class DagCholeskyWorker:
    """DagCholeskyWorker (W2:93-565) on one B200."""

    def __init__(self, device: int | torch.device | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("DagCholeskyWorker needs a CUDA device: the tile ops have no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                   (device if isinstance(device, int) else device.index or 0))
        self._pinned: dict = {}
        self._dev: dict = {}

    # ---- staging ---------------------------------------------------------------------------------
    def _buffers(self, B: int, slot: int):
        key = (B, slot)
        if key not in self._pinned:
            self._pinned[key] = torch.empty((B, B), dtype=torch.float64).pin_memory()
            self._dev[key] = torch.empty((B, B), dtype=torch.float64, device=self.device)
        return self._pinned[key], self._dev[key]

    def _upload(self, blob: bytes, B: int, slot: int) -> torch.Tensor:
        pin, dev = self._buffers(B, slot)
        pin.view(-1).numpy()[:] = np.frombuffer(blob, dtype="<f8")
        dev.copy_(pin, non_blocking=True)
        return dev

    def _download(self, dev: torch.Tensor, B: int, slot: int) -> bytes:
        pin, _ = self._buffers(B, slot)
        pin.copy_(dev, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return pin.numpy().tobytes()

    # ---- Execute ---------------------------------------------------------------------------------
    def Execute(self, taskHandler: TaskHandler) -> ProcessStatus:
        try:
            p = handle_json(taskHandler.getPayload())
            out_id = taskHandler.getExpectedResults()[0]
            deps = taskHandler.getDataDependencies()
            if p.op == "POTRF":
                names = [p.in_]
            elif p.op == "TRSM":
                names = [p.inL, p.inA]
            elif p.op == "SYRK":
                names = [p.inC, p.inA]
            elif p.op == "GEMM":
                names = [p.inC, p.inAi, p.inAj]
            else:
                return ProcessStatus.Error("Unknown op=" + p.op)
            tag = _TAG[p.op]
            tiles = []
            with torch.cuda.device(self.device):
                for slot, name in enumerate(names):
                    if name not in deps:
                        prefix = _MISSING_PREFIX.get((p.op, slot), tag + " Missing dependency: ")
                        return ProcessStatus.Error(prefix + name)
                    blob = deps[name]
                    if p.B <= 0 or len(blob) // 8 != p.B * p.B:
                        return ProcessStatus.Error(f"{tag} Bad block size: expected {p.B * p.B} doubles, got "
                                                   f"{len(blob) // 8}")
                    tiles.append(self._upload(blob, p.B, slot))
                if p.op == "POTRF":
                    info = int(tile_ops.potrf_tile(tiles[0]).item())
                    if info != 0:
                        raise RuntimeError(f"{tag} dpotrf info={info}")
                    out = tiles[0]
                elif p.op == "TRSM":
                    tile_ops.trsm_tile(tiles[0], tiles[1])
                    out = tiles[1]
                elif p.op == "SYRK":
                    tile_ops.syrk_tile(tiles[1], tiles[0])
                    out = tiles[0]
                else:
                    tile_ops.gemm_tile(tiles[1], tiles[2], tiles[0])
                    out = tiles[0]
                data = self._download(out, p.B, 0)
            try:
                taskHandler.send_result(out_id, data)
            except Exception as e:  # noqa: BLE001  (W2:263-266)
                return ProcessStatus.Error(f"{tag} send_result failed: {e}" if p.op == "POTRF"
                                           else f"send_result failed: {e}")
            return ProcessStatus.Ok()
        except Exception as e:  # noqa: BLE001  (W2:558-560)
            return ProcessStatus.Error("Exception: " + _what(e))


def _what(e: Exception) -> str:
    if isinstance(e, KeyError) and e.args:
        return str(e.args[0])
    return str(e)


class WorkerError(RuntimeError):
    """Raised by execute() when the task's ProcessStatus is an error."""


_default_worker: DagCholeskyWorker | None = None


def execute(payload_json: str, deps: Mapping[str, bytes]) -> bytes:
    """One tile task, functional form: returns the output blob or raises WorkerError with the
    reference's status message."""
    global _default_worker
    if _default_worker is None:
        _default_worker = DagCholeskyWorker()
    th = TaskHandler(payload_json, deps)
    st = _default_worker.Execute(th)
    if not st.ok:
        raise WorkerError(st.details)
    return th.results[th.expected_results[0]]
