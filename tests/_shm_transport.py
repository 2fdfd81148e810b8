"""TEST DOUBLE (tests/ only): the peer-push panel transport (dense_linear_app_b200/transport.py) with
its four primitives played on the CPU — the "peer-mapped" buffers are POSIX shared memory between the
gloo ranks, pushes are memmoves, flag waits are host polls with a timeout (a wait that never ends is
reported as a deadlock).  Everything above the primitives — slots, subsets, offsets, credits, flag
values — is the product's own code, so the multi-process CPU tests check exactly what the GPUs run."""
import ctypes as C
import os
import time
from multiprocessing import shared_memory

import torch.distributed as dist

from dense_linear_app_b200.transport import PeerTransport

TIMEOUT_S = 60.0


class ShmTransport(PeerTransport):
    def _open(self):
        tag = os.environ.get("MASTER_PORT", "0")
        self._shm = shared_memory.SharedMemory(create=True, size=self.nbytes, name=f"chol_{tag}_{id(self) & 0xffff}_{self.rank}")
        self._shm.buf[:] = bytes(self.nbytes)
        self.local = C.addressof(C.c_char.from_buffer(self._shm.buf))
        names = [None] * self.world
        dist.all_gather_object(names, self._shm.name, group=self.group)
        self._peers = {}
        for r, n in enumerate(names):
            if r != self.rank:
                m = shared_memory.SharedMemory(name=n)
                self._peers[r] = m
                self.peer_base[r] = C.addressof(C.c_char.from_buffer(m.buf))
        dist.barrier(group=self.group)
        self.pushed_bytes = 0

    def close(self):
        dist.barrier(group=self.group)
        self.peer_base, self.local = {}, 0
        # ctypes views pin the buffers: drop the mappings without closing (process exit unmaps)
        dist.barrier(group=self.group)
        try:
            self._shm.unlink()
        except FileNotFoundError:
            pass

    @staticmethod
    def _word(addr):
        return C.c_uint32.from_address(addr)

    def _poll(self, addr, value, what):
        t0 = time.time()
        while True:
            d = (self._word(addr).value - value) & 0xFFFFFFFF
            if d < 0x80000000:
                return
            if time.time() - t0 > TIMEOUT_S:
                raise RuntimeError(f"rank {self.rank}: {what} never reached {value} (deadlock)")
            time.sleep(0.0002)

    def _send(self, pushes, ready_stream, send_stream_of):
        for p in pushes:
            if p.credit is not None:
                self._poll(self.local + 4 * p.credit, p.credit_value, f"credit word {p.credit}")
            base = self.peer_base[p.peer]
            for t in range(p.count):
                C.memmove(base + p.dst_off + t * p.dst_stride * p.tile_bytes, p.src + t * p.src_stride * p.tile_bytes,
                          p.tile_bytes)
            self.pushed_bytes += p.count * p.tile_bytes
            if p.flag is not None:
                self._word(base + 4 * p.flag).value = p.flag_value

    def _wait(self, flag, value, stream):
        self._poll(self.local + 4 * flag, value, f"data word {flag}")

    def _post(self, targets, value, stream):
        for r, f in targets:
            self._word(self.peer_base[r] + 4 * f).value = value
