// aux.cuh — generators, norms and copies around the factorization (off the timed path),
// plus the FP64 peak microbenchmarks that give the roofline denominators.
//   V6:46  CHAMELEON_dplgsy_Tile(bump, ChamLower, descA, seed)   -> plgsy_tile_kernel
//   V6:72  CHAMELEON_dlange_Tile(ChamInfNorm, desc)              -> tile_abs_sums kernels
//   V6:77  CHAMELEON_dlacpy_Tile(ChamLower, descA, descR)        -> tile_tril_kernel
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace chol {

// ---- dplgsy-style LCG (constants as in Chameleon's coreblas random.h, recalled; the
// generator is counter based: element (i,j), i>=j, uses LCG state number i + j*bigM) ----
constexpr unsigned long long RND64_A = 6364136223846793005ULL;
constexpr unsigned long long RND64_C = 1ULL;
constexpr double RNDF_MUL = 5.4210108624275222e-20;

__host__ __device__ inline unsigned long long rnd64_jump(unsigned long long n, unsigned long long seed) {
    unsigned long long a_k = RND64_A, c_k = RND64_C, ran = seed;
    for (; n; n >>= 1) {
        if (n & 1) ran = a_k * ran + c_k;
        c_k *= (a_k + 1);
        a_k *= a_k;
    }
    return ran;
}

__device__ __forceinline__ double rnd64_value(unsigned long long ran) {
    // 0.5f - ran * RndF_Mul, evaluated with two separately rounded operations so that the
    // C oracle (-ffp-contract=off) and numpy produce the same bits.
    return __dsub_rn(0.5, __dmul_rn(__ull2double_rn(ran), RNDF_MUL));
}

constexpr int PLGSY_ROWS_PER_THREAD = 4;

__global__ void plgsy_tile_kernel(double bump, int mb, int nb, double* __restrict__ A, int lda,
                                  unsigned long long bigM, long long row0, long long col0, long long N,
                                  unsigned long long seed) {
    const int c = blockIdx.y;
    const int r_begin = (blockIdx.x * blockDim.x + threadIdx.x) * PLGSY_ROWS_PER_THREAD;
    if (c >= nb || r_begin >= mb) return;
    const long long gj = col0 + c;
    unsigned long long ran = 0;
    bool have = false;  // ran holds the state of the previous row in the same (lower) column walk
#pragma unroll
    for (int u = 0; u < PLGSY_ROWS_PER_THREAD; ++u) {
        const int r = r_begin + u;
        if (r >= mb) break;
        const long long gi = row0 + r;
        double v;
        if (gi >= N || gj >= N) {
            v = (gi == gj) ? 1.0 : 0.0;  // identity padding of ragged edge tiles
            have = false;
        } else if (gi >= gj) {
            if (have) ran = RND64_A * ran + RND64_C;
            else ran = rnd64_jump((unsigned long long)gi + (unsigned long long)gj * bigM, seed);
            have = true;
            v = rnd64_value(ran);
            if (gi == gj) v += bump;
        } else {
            // strict upper part: mirror of element (gj, gi)
            const unsigned long long rr = rnd64_jump((unsigned long long)gj + (unsigned long long)gi * bigM, seed);
            v = rnd64_value(rr);
            have = false;
        }
        A[size_t(c) * lda + r] = v;
    }
}

// ---- per-column sum of squares; one warp per column --------------------------------
// mode 0: all rows; mode 1: lower triangle as a symmetric matrix (strict lower counted
// twice, diagonal once, rows above the diagonal ignored).
__global__ void tile_col_sumsq_kernel(int m, int n, const double* __restrict__ A, int lda, int mode,
                                      double* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n) return;
    const int j = warp;
    const double* col = A + size_t(j) * lda;
    double s = 0.0;
    const int i0 = mode ? j : 0;
    for (int i = i0 + lane; i < m; i += 32) {
        const double v = col[i];
        const double w = (mode && i > j) ? 2.0 : 1.0;
        s = fma(w * v, v, s);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[j] = s;
}

// ---- |a| column sums (one warp per column) and row sums (one thread per row) ---------
__global__ void tile_abs_colsum_kernel(int m, int n, const double* __restrict__ A, int lda, int mode,
                                       double* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n) return;
    const int j = warp;
    const double* col = A + size_t(j) * lda;
    double s = 0.0;
    for (int i = (mode ? j : 0) + lane; i < m; i += 32) s += fabs(col[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[j] = s;
}
__global__ void tile_abs_rowsum_kernel(int m, int n, const double* __restrict__ A, int lda, int mode,
                                       double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double s = 0.0;
    const int jend = mode ? min(n, i + 1) : n;
    for (int j = 0; j < jend; ++j) s += fabs(A[size_t(j) * lda + i]);
    out[i] = s;
}

__global__ void tile_tril_kernel(int n, const double* __restrict__ A, int lda, double* __restrict__ B, int ldb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= n || j >= n) return;
    B[size_t(j) * ldb + i] = (i >= j) ? A[size_t(j) * lda + i] : 0.0;
}

// In-place transposition of `ntiles` n x n column-major tiles (tile t at A + t * stride): the bridge between
// uplo = Upper and the lower-triangular kernels (A = U^T U  <=>  A = L L^T with L = U^T).  One CTA per pair
// of 32 x 32 sub-blocks (I, J), I >= J, of one tile; both are staged in shared memory and written back swapped.
__global__ void __launch_bounds__(256) tile_transpose_kernel(int n, double* __restrict__ A, int lda, long long stride) {
    __shared__ double sa[32][33], sb[32][33];
    const int nb = (n + 31) / 32;
    int I = 0;
    while ((I + 1) * (I + 2) / 2 <= int(blockIdx.x)) ++I;
    const int J = int(blockIdx.x) - I * (I + 1) / 2;
    if (I >= nb) return;
    double* T = A + size_t(blockIdx.y) * size_t(stride);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    for (int c = ty; c < 32; c += 8) {
        const int ra = I * 32 + tx, ca = J * 32 + c;             // block (I, J): rows I, columns J
        if (ra < n && ca < n) sa[c][tx] = T[size_t(ca) * lda + ra];
        const int rb = J * 32 + tx, cb = I * 32 + c;             // block (J, I)
        if (I != J && rb < n && cb < n) sb[c][tx] = T[size_t(cb) * lda + rb];
    }
    __syncthreads();
    for (int c = ty; c < 32; c += 8) {
        // new (J, I)[r][c'] = old (I, J)[c'][r]
        const int rb = J * 32 + tx, cb = I * 32 + c;
        if (rb < n && cb < n) T[size_t(cb) * lda + rb] = sa[tx][c];
        const int ra = I * 32 + tx, ca = J * 32 + c;
        if (I != J && ra < n && ca < n) T[size_t(ca) * lda + ra] = sb[tx][c];
    }
}

// ---- FP64 peak microbenchmarks -------------------------------------------------------
// kind 0: 16 independent DFMA chains per thread.
__global__ void __launch_bounds__(1024, 1) peak_dfma_kernel(int iters, double* out) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double a = 0.999999, b = 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}
// kind 1..3: 32 independent m8n8k4 DMMA accumulators per warp (same ILP as the GEMM).
__global__ void __launch_bounds__(256, 1) peak_dmma_kernel(int iters, double* out) {
    double acc[32][2];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1e-3 * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (threadIdx.x - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc[i * 4 + j][0]), "+d"(acc[i * 4 + j][1])
                             : "d"(a[i]), "d"(b[j]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) out[0] = s;
}

}  // namespace chol
