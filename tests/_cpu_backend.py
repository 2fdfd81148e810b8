"""TEST DOUBLE (tests/ only): the TiledCholesky schedule with its three kernel entry points
executed by the CPU oracle on host tensors, so the host logic — plan, task pointers, block-cyclic
ownership, broadcast order — is testable without a GPU (gloo, world_size > 1).  Not part of the
product: the package itself has no CPU path."""
import ctypes as C

import numpy as np

from dense_linear_app_b200.cholesky import TiledCholesky
from oracle import oracle as O


class OracleBackedCholesky(TiledCholesky):
    def _potrf_workspace(self, b):
        return 8

    def _k_potrf(self, a_ptr, info_base, st):
        info = O.lib().oracle_potrf_tile(self.b, a_ptr, self.b)
        if info and int(self.d_info[0]) == 0:
            self.d_info[0] = info_base + info

    def _make_transport(self):
        from _shm_transport import ShmTransport
        return ShmTransport(self.nt, self.b, self.grid, self.rank, self.work.numel(), self.group)

    def _k_trsm_panel(self, l_ptr, work_ptr, tiles_ptr, ntiles, st):
        ptrs = (C.c_int64 * ntiles).from_address(tiles_ptr)
        for t in range(ntiles):
            O.lib().oracle_trsm_tile(self.b, self.b, l_ptr, self.b, ptrs[t], self.b)

    def _k_update(self, tasks_ptr, ntasks, st, thin=False):
        rec = np.ctypeslib.as_array((C.c_int64 * (4 * ntasks)).from_address(tasks_ptr)).reshape(ntasks, 4)
        b = self.b
        for c, a, bb, flag in rec.tolist():
            if flag & 1:
                assert a == bb or True
                # lower-only update  C -= A B^T  (A == B for a SYRK; the residual's T_k T_k^T too)
                if a == bb:
                    O.lib().oracle_syrk_tile(b, b, a, b, c, b)
                else:
                    _lower_gemm(b, a, bb, c)
            else:
                O.lib().oracle_gemm_tile(b, b, b, a, b, bb, b, c, b)

    def _k_tril(self, src_ptr, dst_ptr, st):
        b = self.b
        src = np.ctypeslib.as_array((C.c_double * (b * b)).from_address(src_ptr)).reshape(b, b)  # [c, r]
        dst = np.ctypeslib.as_array((C.c_double * (b * b)).from_address(dst_ptr)).reshape(b, b)
        dst[...] = np.triu(src)  # lower triangle of the column-major tile


def _lower_gemm(b, a, bb, c):
    A = np.ctypeslib.as_array((C.c_double * (b * b)).from_address(a)).reshape(b, b).T
    B = np.ctypeslib.as_array((C.c_double * (b * b)).from_address(bb)).reshape(b, b).T
    Cm = np.ctypeslib.as_array((C.c_double * (b * b)).from_address(c)).reshape(b, b)
    upd = np.tril(A @ B.T)
    Cm -= upd.T
