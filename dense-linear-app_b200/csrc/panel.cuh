// panel.cuh — the latency-bound pieces of the factorization: Cholesky of one diagonal
// block (n <= 128) resident in shared memory, and inversion of that triangular block.
//
// POTRF tile op (W2:238 CHAMELEON_dpotrf_Tile(ChamLower, dA)) is built in chol_abi.cu as a
// blocked right-looking sweep with block size NBD=128: diag block here, the rest as
// DMMA rank-128 updates (gemm_dmma.cuh).  TRSM (W2:323) uses the inverted diagonal blocks
// produced here, so it also runs on the DMMA kernel.
#pragma once
#include <cuda_runtime.h>

namespace chol {

constexpr int NBD = 128;            // diagonal block size of the blocked POTRF / TRSM
constexpr int DPITCH = NBD + 1;     // smem pitch (doubles): odd -> row and column walks conflict-free
constexpr int DIAG_THREADS = 512;
constexpr size_t DIAG_SMEM_BYTES = size_t(NBD) * DPITCH * 8 + 64;

// In-place lower Cholesky of the n x n block held column-major in S (pitch DPITCH).
// Right-looking, one column per step; returns (to every thread) 0 or the 1-based index of
// the first non-positive pivot (LAPACK dpotrf info).
__device__ __forceinline__ int potrf_smem(double* S, int n, int* s_info) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) *s_info = 0;
    __syncthreads();
    for (int c = 0; c < n; ++c) {
        // pivot
        const double d = S[c * DPITCH + c];
        if (!(d > 0.0)) {  // also catches NaN
            if (tid == 0 && *s_info == 0) *s_info = c + 1;
        }
        const double piv = sqrt(d);
        const double inv = 1.0 / piv;
        __syncthreads();  // everyone has read S[c][c]
        // scale column c (every thread scales the entries it will need? no: cooperative)
        for (int i = c + tid; i < n; i += nt) {
            S[c * DPITCH + i] = (i == c) ? piv : S[c * DPITCH + i] * inv;
        }
        __syncthreads();
        // rank-1 update of the trailing lower triangle: S[i][j] -= l_i * l_j, c < j <= i < n
        const int rem = n - c - 1;
        // map a linear index over the rem x rem square, skip the upper part (cheap, rem <= 127)
        for (int idx = tid; idx < rem * rem; idx += nt) {
            const int jj = idx / rem, ii = idx - jj * rem;
            if (ii >= jj) {
                const int i = c + 1 + ii, j = c + 1 + jj;
                S[j * DPITCH + i] = fma(-S[c * DPITCH + i], S[c * DPITCH + j], S[j * DPITCH + i]);
            }
        }
        __syncthreads();
    }
    return *s_info;
}

// In-place inversion of the lower-triangular n x n block in S (LAPACK dtrti2, lower,
// non-unit, processed from the last column to the first):
//   W[j][j] = 1/L[j][j];  W[j+1:, j] = -W[j+1:, j+1:] * L[j+1:, j] * W[j][j]
// `x` is an n-vector of scratch.  Row i of the matrix-vector product is computed by a
// group of 4 threads (dot product split 4 ways, combined with shuffles).
__device__ __forceinline__ void trtri_smem(double* S, int n, double* x) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int grp = tid >> 2, sub = tid & 3;
    static_assert(DIAG_THREADS / 4 >= NBD, "one 4-thread group per row");
    for (int j = n - 1; j >= 0; --j) {
        const double wjj = 1.0 / S[j * DPITCH + j];
        __syncthreads();
        // x = L[j+1:, j]
        for (int i = j + 1 + tid; i < n; i += nt) x[i] = S[j * DPITCH + i];
        __syncthreads();
        // y_i = sum_{k=j+1..i} W[i][k] * x[k]
        // (ngrp >= NBD, so one row per group; the shuffles are executed by every lane)
        {
            const int i = j + 1 + grp;
            double s = 0.0;
            if (i < n)
                for (int k = j + 1 + sub; k <= i; k += 4) s = fma(S[k * DPITCH + i], x[k], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (i < n && sub == 0) S[j * DPITCH + i] = -s * wjj;
        }
        if (tid == 0) S[j * DPITCH + j] = wjj;
        __syncthreads();
    }
}

// Factor the n x n diagonal block at A (col-major, lda) in place (lower; strict upper left
// untouched) and write the inverse of the factor (full n x n, upper part zero, ld = NBD)
// to Winv.  One CTA.  info: first failure wins (device-wide, stream ordered).
__global__ void __launch_bounds__(DIAG_THREADS, 1)
potrf_diag_kernel(int n, double* __restrict__ A, int lda, double* __restrict__ Winv, int* d_info, int info_base) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* S = reinterpret_cast<double*>(smem_raw);
    __shared__ double xs[NBD];
    __shared__ int s_info;
    const int tid = threadIdx.x, nt = blockDim.x;
    // load lower triangle (coalesced down the columns)
    for (int idx = tid; idx < n * n; idx += nt) {
        const int j = idx / n, i = idx - j * n;
        S[j * DPITCH + i] = (i >= j) ? A[size_t(j) * lda + i] : 0.0;
    }
    __syncthreads();
    const int info = potrf_smem(S, n, &s_info);
    if (info != 0 && tid == 0 && d_info) atomicCAS(d_info, 0, info_base + info);
    // store L (lower triangle only)
    for (int idx = tid; idx < n * n; idx += nt) {
        const int j = idx / n, i = idx - j * n;
        if (i >= j) A[size_t(j) * lda + i] = S[j * DPITCH + i];
    }
    __syncthreads();
    trtri_smem(S, n, xs);
    for (int idx = tid; idx < n * n; idx += nt) {
        const int j = idx / n, i = idx - j * n;
        Winv[size_t(j) * NBD + i] = (i >= j) ? S[j * DPITCH + i] : 0.0;
    }
}

// Invert the nblk diagonal blocks (NBD x NBD, last one possibly smaller) of the lower
// triangular b x b matrix L: block q -> Winv + q*NBD*NBD (ld = NBD, upper part zero).
// One CTA per block.  Used by the stateless TRSM tile op, where only L arrives (W2:273-323).
__global__ void __launch_bounds__(DIAG_THREADS, 1)
trtri_diag_kernel(int b, const double* __restrict__ L, int ldl, double* __restrict__ Winv) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* S = reinterpret_cast<double*>(smem_raw);
    __shared__ double xs[NBD];
    const int q = blockIdx.x;
    const int o = q * NBD;
    const int n = min(NBD, b - o);
    const int tid = threadIdx.x, nt = blockDim.x;
    const double* A = L + size_t(o) * ldl + o;
    for (int idx = tid; idx < n * n; idx += nt) {
        const int j = idx / n, i = idx - j * n;
        S[j * DPITCH + i] = (i >= j) ? A[size_t(j) * ldl + i] : 0.0;
    }
    __syncthreads();
    trtri_smem(S, n, xs);
    double* W = Winv + size_t(q) * NBD * NBD;
    for (int idx = tid; idx < n * n; idx += nt) {
        const int j = idx / n, i = idx - j * n;
        W[size_t(j) * NBD + i] = (i >= j) ? S[j * DPITCH + i] : 0.0;
    }
}

}  // namespace chol
