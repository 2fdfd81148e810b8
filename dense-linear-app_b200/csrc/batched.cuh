// batched.cuh — many independent small Cholesky factorizations (the ArmoniK many-task
// workload: one POTRF task per small tile, C1:139-141 / W2:179-268, here one CTA per matrix).
#pragma once
#include <cuda_runtime.h>
#include "panel.cuh"

namespace chol {

// n <= 128: whole matrix resident in shared memory, factored with the warp-level recursive
// block Cholesky of panel.cuh.
__global__ void __launch_bounds__(DIAG_THREADS, 1)
potrf_batched_smem_kernel(int n, double* __restrict__ Abase, int lda, long long stride, int* __restrict__ d_info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_info;
    const DiagSmem m = diag_smem(smem_raw);
    double* A = Abase + size_t(blockIdx.x) * stride;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int n32 = (n + SB - 1) / SB * SB;
    diag_load(m.S, A, lda, n, n32);
    __syncthreads();
    const int info = potrf_block_smem(m, n32, &s_info);
    if (tid == 0) d_info[blockIdx.x] = info;
    for (int idx = tid; idx < n * n; idx += nt) {
        const int j = idx / n, i = idx - j * n;
        if (i >= j) A[size_t(j) * lda + i] = m.S[j * DPITCH + i];
    }
}

// n > 128: blocked right-looking with the current block column staged in shared memory.
//   for each block column (width W=32):
//     load the panel (rows o..n-1, W columns) into smem, factor its top WxW block and solve the
//     rows below inside smem, write the panel back, then update the trailing lower triangle
//     in global memory (L2 resident: one matrix is <= a few hundred KB) from the smem panel.
constexpr int BATCHED_GLOBAL_THREADS = 512;
constexpr int BW = 32;          // block-column width
constexpr int BPITCH = BW + 1;  // panel stored row-major in smem: P[i][c] at i*BPITCH + c
constexpr int BATCHED_MAX_N = 880;  // n * BPITCH * 8 bytes of dynamic smem must fit 227 KB
inline size_t batched_global_smem(int n) { return size_t(n) * BPITCH * 8; }

__global__ void __launch_bounds__(BATCHED_GLOBAL_THREADS)
potrf_batched_global_kernel(int n, double* __restrict__ Abase, int lda, long long stride, int* __restrict__ d_info) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    __shared__ int s_info;
    double* P = reinterpret_cast<double*>(smem_dyn);  // (n - o) x BW panel, row-major, pitch BPITCH
    double* A = Abase + size_t(blockIdx.x) * stride;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) s_info = 0;
    __syncthreads();
    for (int o = 0; o < n; o += BW) {
        const int w = min(BW, n - o);
        const int rows = n - o;
        // load panel
        for (int idx = tid; idx < rows * w; idx += nt) {
            const int c = idx / rows, i = idx - c * rows;
            P[i * BPITCH + c] = A[size_t(o + c) * lda + (o + i)];
        }
        __syncthreads();
        // factor the panel column by column (unblocked, all rows at once)
        for (int c = 0; c < w; ++c) {
            const double d = P[c * BPITCH + c];
            if (!(d > 0.0) && tid == 0 && s_info == 0) s_info = o + c + 1;
            const double piv = sqrt(d), inv = 1.0 / piv;
            __syncthreads();
            for (int i = c + tid; i < rows; i += nt) P[i * BPITCH + c] = (i == c) ? piv : P[i * BPITCH + c] * inv;
            __syncthreads();
            // update the remaining panel columns c+1..w-1 for rows >= that column
            const int rc = w - c - 1;
            for (int idx = tid; idx < rows * rc; idx += nt) {
                const int i = idx / rc, cc = c + 1 + (idx - i * rc);
                if (i >= cc) P[i * BPITCH + cc] = fma(-P[i * BPITCH + c], P[cc * BPITCH + c], P[i * BPITCH + cc]);
            }
            __syncthreads();
        }
        // write the panel back (lower part only)
        for (int idx = tid; idx < rows * w; idx += nt) {
            const int c = idx / rows, i = idx - c * rows;
            if (i >= c) A[size_t(o + c) * lda + (o + i)] = P[i * BPITCH + c];
        }
        // trailing update: A[i][j] -= sum_c P[i][c] * P[j][c], o+w <= j <= i < n
        const int rem = rows - w;
        for (int idx = tid; idx < rem * rem; idx += nt) {
            const int jj = idx / rem, ii = idx - jj * rem;
            if (ii >= jj) {
                const double* pi = P + (w + ii) * BPITCH;
                const double* pj = P + (w + jj) * BPITCH;
                double s = 0.0;
                for (int c = 0; c < w; ++c) s = fma(pi[c], pj[c], s);
                A[size_t(o + w + jj) * lda + (o + w + ii)] -= s;
            }
        }
        __syncthreads();
    }
    if (tid == 0) d_info[blockIdx.x] = s_info;
}

}  // namespace chol
