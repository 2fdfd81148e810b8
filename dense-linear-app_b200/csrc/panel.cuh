// panel.cuh — the latency-bound pieces of the factorization: Cholesky of one diagonal block
// (n <= 128) resident in shared memory, and the inverse of its triangular factor.
//
// POTRF tile op (W2:238 CHAMELEON_dpotrf_Tile(ChamLower, dA)) is built in chol_abi.cu as a
// blocked right-looking sweep with block size NBD=128: diagonal block here, the rest as DMMA
// rank-128 updates (gemm_dmma.cuh).  TRSM (W2:323) multiplies by the inverted diagonal blocks
// produced here, so it runs on the DMMA kernel too.
//
// The 128x128 block is itself factored recursively with 32x32 sub-blocks:
//   * potrf32_warp : one warp, one matrix row per lane held in 32 registers; column
//     elimination with warp shuffles (pivot broadcast + l_jc broadcast), no shared-memory
//     traffic and no block barrier inside the 32 column steps;
//   * trtri32_warp : one warp, lane j runs the forward substitution for column j of inv(L)
//     in axpy form (32 independent FMA chains per step, L read as shared-memory broadcasts);
//   * the 32-wide panel below a sub-block is solved by multiplying with that inverse, the
//     trailing part is updated with 4x4 register micro-tiles (all 512 threads);
//   * the inverse of the whole block is assembled level by level from the four 32x32 inverses
//     (W_IJ = -D_I * sum_K L_IK W_KJ) and stored transposed in the unused upper part of S.
// Blocks smaller than a multiple of 32 are padded with the identity, chol(diag(A, I)) =
// diag(chol(A), I), so no edge logic exists past load/store.
#pragma once
#include <cuda_runtime.h>

namespace chol {

constexpr int NBD = 128;            // diagonal block size of the blocked POTRF / TRSM
constexpr int DPITCH = NBD + 1;     // smem pitch (doubles): odd -> row and column walks conflict-free
constexpr int DIAG_THREADS = 512;
constexpr int SB = 32;              // sub-block (one warp)
constexpr int SBP = SB + 1;         // pitch of the 32x32 scratch blocks
constexpr int DIAG_S_DOUBLES = NBD * DPITCH;
constexpr int DIAG_D_DOUBLES = 4 * SB * SBP;   // inverses of the four diagonal sub-blocks
constexpr int DIAG_T_DOUBLES = 3 * SB * SBP;   // products of one level of the inverse assembly
constexpr size_t DIAG_SMEM_BYTES = size_t(DIAG_S_DOUBLES + DIAG_D_DOUBLES + DIAG_T_DOUBLES + NBD) * 8 + 64;

struct DiagSmem {
    double* S;     // NBD x NBD, column-major, pitch DPITCH: the block (lower) / W off-diagonal (upper, transposed)
    double* D;     // D[J] = inv(L_JJ), 32 x 32 column-major pitch SBP, upper part zero
    double* T;     // scratch
    double* invd;  // 1 / l_cc
};
__device__ __forceinline__ DiagSmem diag_smem(unsigned char* raw) {
    DiagSmem m;
    m.S = reinterpret_cast<double*>(raw);
    m.D = m.S + DIAG_S_DOUBLES;
    m.T = m.D + DIAG_D_DOUBLES;
    m.invd = m.T + DIAG_T_DOUBLES;
    return m;
}

// Load the lower triangle of the n x n block at A into S, padded to n32 = roundup(n, 32) with the
// identity.  Thread t owns row i = t % n32 and every (blockDim/n32)-th column; the global loads
// of 8 columns are issued back to back before the first store (ncu: the straightforward
// one-element-per-iteration loop spent 15 % of the kernel waiting on serialized load latency).
__device__ __forceinline__ void diag_load(double* S, const double* __restrict__ A, int lda, int n, int n32) {
    const int i = threadIdx.x % n32;
    const int cstep = blockDim.x / n32;          // 4 (n32 = 128) .. 16 (n32 = 32)
    const int jfirst = threadIdx.x / n32;
    if (jfirst >= cstep) return;                 // leftover threads when n32 does not divide blockDim
    for (int j0 = jfirst; j0 < n32; j0 += 8 * cstep) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + u * cstep;
            v[u] = (j < n32 && i == j) ? 1.0 : 0.0;
            if (i < n && j < n && i >= j) v[u] = A[size_t(j) * lda + i];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + u * cstep;
            if (j < n32) S[j * DPITCH + i] = v[u];
        }
    }
}

// One warp: in-place Cholesky of the 32x32 sub-block at offset o.  Lane i holds row i.
// Returns 0 or the 1-based local index of the first non-positive pivot (also for NaN).
__device__ __forceinline__ int potrf32_warp(double* S, int o, double* invd) {
    const int lane = threadIdx.x & 31;
    double a[SB];
#pragma unroll
    for (int j = 0; j < SB; ++j) a[j] = S[(o + j) * DPITCH + o + lane];
    int info = 0;
#pragma unroll
    for (int c = 0; c < SB; ++c) {
        const double d = __shfl_sync(0xffffffffu, a[c], c);
        if (!(d > 0.0) && info == 0) info = c + 1;
        const double piv = sqrt(d);
        const double inv = 1.0 / piv;
        const double l = (lane == c) ? piv : a[c] * inv;
        a[c] = l;
        if (lane == c) invd[o + c] = inv;
#pragma unroll
        for (int j = c + 1; j < SB; ++j) {
            const double ljc = __shfl_sync(0xffffffffu, l, j);
            a[j] = fma(-l, ljc, a[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < SB; ++j)
        if (lane >= j) S[(o + j) * DPITCH + o + lane] = a[j];
    return info;
}

// One warp: Dj = inverse of the lower-triangular 32x32 sub-block at offset o (reads S and invd).
// Lane j computes column j by forward substitution, axpy form.
__device__ __forceinline__ void trtri32_warp(const double* S, int o, const double* invd, double* Dj) {
    const int lane = threadIdx.x & 31;
    double acc[SB];
#pragma unroll
    for (int i = 0; i < SB; ++i) acc[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < SB; ++k) {
        const double w = acc[k] * invd[o + k];
        acc[k] = w;
#pragma unroll
        for (int i = k + 1; i < SB; ++i) acc[i] = fma(-S[(o + k) * DPITCH + o + i], w, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < SB; ++i) Dj[lane * SBP + i] = acc[i];
}

// triangular index -> (ti, tj), tj <= ti, t = ti*(ti+1)/2 + tj
__device__ __forceinline__ void tri_decode(int t, int& ti, int& tj) {
    int r = int((sqrtf(8.0f * float(t) + 1.0f) - 1.0f) * 0.5f);
    while (r * (r + 1) / 2 > t) --r;
    while ((r + 1) * (r + 2) / 2 <= t) ++r;
    ti = r;
    tj = t - r * (r + 1) / 2;
}

// In-place lower Cholesky of the n32 x n32 block in S (n32 multiple of 32, <= 128) by the whole
// CTA; also leaves D[J] = inv(L_JJ) for every sub-block.  Returns LAPACK info to every thread.
__device__ __forceinline__ int potrf_block_smem(const DiagSmem& m, int n32, int* s_info) {
    const int tid = threadIdx.x;
    // warp index obtained through a shuffle: tells the compiler it is warp-uniform, so the
    // `warp == 0` branch below is convergent and the shuffles inside it are emitted bare
    // (without it ptxas wraps each one in WARPSYNC/ENDCOLLECTIVE: 30k vs 13k SASS lines)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    double* S = m.S;
    if (tid == 0) *s_info = 0;
    __syncthreads();
    for (int o = 0; o < n32; o += SB) {
        double* Dj = m.D + (o / SB) * SB * SBP;
        if (warp == 0) {
            const int info = potrf32_warp(S, o, m.invd);
            if (info != 0 && (tid & 31) == 0 && *s_info == 0) *s_info = o + info;
            __syncwarp();
            trtri32_warp(S, o, m.invd, Dj);
        }
        __syncthreads();
        const int R = n32 - o - SB;   // rows below this sub-block: 0, 32, 64 or 96
        if (R == 0) break;
        // ---- panel: X = A_panel * Dj^T, X(r,c) = sum_{k<=c} A(r,k) Dj(c,k); 4 threads per row
        double out[8];
        const int cg = tid / R;       // warp-uniform (R is a multiple of 32)
        const int r = o + SB + (tid - cg * R);
        const bool act = tid < 4 * R;
        if (act) {
            double x[SB];
#pragma unroll
            for (int k = 0; k < SB; ++k) x[k] = (k < 8 * (cg + 1)) ? S[(o + k) * DPITCH + r] : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = 8 * cg + u;
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < SB; ++k)
                    if (k <= c) s = fma(x[k], Dj[k * SBP + c], s);
                out[u] = s;
            }
        }
        __syncthreads();
        if (act) {
#pragma unroll
            for (int u = 0; u < 8; ++u) S[(o + 8 * cg + u) * DPITCH + r] = out[u];
        }
        __syncthreads();
        // ---- trailing update of the lower triangle, 4x4 micro-tiles: S(i,j) -= sum_k X(i,k) X(j,k)
        const int mt = R / 4;
        const int ntile = mt * (mt + 1) / 2;
        if (tid < ntile) {
            int ti, tj;
            tri_decode(tid, ti, tj);
            const int i0 = o + SB + 4 * ti, j0 = o + SB + 4 * tj;
            double c[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) c[a][b] = 0.0;
#pragma unroll 8
            for (int k = 0; k < SB; ++k) {
                const double* col = S + (o + k) * DPITCH;
                double xi[4], xj[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) { xi[a] = col[i0 + a]; xj[a] = col[j0 + a]; }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) c[a][b] = fma(xi[a], xj[b], c[a][b]);
            }
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    if (i0 + a >= j0 + b) S[(j0 + b) * DPITCH + i0 + a] -= c[a][b];
        }
        __syncthreads();
    }
    __syncthreads();
    return *s_info;
}

// Assemble inv(L) of the n32 x n32 lower-triangular block in S from the D[J] (already computed):
// off-diagonal block W_IJ (I > J) is stored TRANSPOSED in the block above the diagonal,
// W_IJ(r, c) at S[(32 I + r) * DPITCH + 32 J + c].  Level d handles the pairs I = J + d.
__device__ __forceinline__ void trtri_assemble_smem(const DiagSmem& m, int n32) {
    const int tid = threadIdx.x;
    double* S = m.S;
    const int nb = n32 / SB;
    for (int d = 1; d < nb; ++d) {
        const int npair = nb - d;
        const int pair = tid >> 6;              // 64 micro-tiles (4x4) per 32x32 product
        const int mtile = tid & 63;
        const int r0 = 4 * (mtile & 7), c0 = 4 * (mtile >> 3);
        const bool act = pair < npair;
        const int J = pair, I = pair + d;
        double* Tp = m.T + pair * SB * SBP;
        if (act) {
            // T(r,c) = sum_{K=J}^{I-1} sum_k L_IK(r,k) W_KJ(k,c)
            double c[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) c[a][b] = 0.0;
            for (int K = J; K < I; ++K) {
                const double* Dk = m.D + K * SB * SBP;
#pragma unroll 4
                for (int k = 0; k < SB; ++k) {
                    double lv[4], wv[4];
                    const double* lcol = S + (SB * K + k) * DPITCH + SB * I + r0;
#pragma unroll
                    for (int a = 0; a < 4; ++a) lv[a] = lcol[a];
                    if (K == J) {
#pragma unroll
                        for (int b = 0; b < 4; ++b) wv[b] = Dk[(c0 + b) * SBP + k];
                    } else {
                        const double* wrow = S + (SB * K + k) * DPITCH + SB * J + c0;
#pragma unroll
                        for (int b = 0; b < 4; ++b) wv[b] = wrow[b];
                    }
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) c[a][b] = fma(lv[a], wv[b], c[a][b]);
                }
            }
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int a = 0; a < 4; ++a) Tp[(c0 + b) * SBP + r0 + a] = c[a][b];
        }
        __syncthreads();
        if (act) {
            // W_IJ(r,c) = -sum_k D_I(r,k) T(k,c)
            const double* Di = m.D + I * SB * SBP;
            double c[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) c[a][b] = 0.0;
#pragma unroll 4
            for (int k = 0; k < SB; ++k) {
                double dv[4], tv[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) dv[a] = Di[k * SBP + r0 + a];
#pragma unroll
                for (int b = 0; b < 4; ++b) tv[b] = Tp[(c0 + b) * SBP + k];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) c[a][b] = fma(dv[a], tv[b], c[a][b]);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) S[(SB * I + r0 + a) * DPITCH + SB * J + c0 + b] = -c[a][b];
        }
        __syncthreads();
    }
}

// Element (i, j), i >= j, of the assembled inverse.
__device__ __forceinline__ double trtri_read(const DiagSmem& m, int i, int j) {
    const int I = i / SB, J = j / SB;
    if (I == J) return m.D[I * SB * SBP + (j - J * SB) * SBP + (i - I * SB)];
    return m.S[i * DPITCH + j];
}

__device__ __forceinline__ void trtri_store(const DiagSmem& m, int n, double* __restrict__ Winv) {
    // full n x n, upper part zero, ld = NBD
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int j = idx / n, i = idx - j * n;
        Winv[size_t(j) * NBD + i] = (i >= j) ? trtri_read(m, i, j) : 0.0;
    }
}

// Inverse of every 32x32 diagonal sub-block of the lower-triangular block in S (one warp each).
__device__ __forceinline__ void trtri_subblocks_smem(const DiagSmem& m, int n32) {
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int idx = tid; idx < n32; idx += blockDim.x) m.invd[idx] = 1.0 / m.S[idx * DPITCH + idx];
    __syncthreads();
    if (warp < n32 / SB) trtri32_warp(m.S, warp * SB, m.invd, m.D + warp * SB * SBP);
    __syncthreads();
}

// Factor the n x n diagonal block at A (col-major, lda) in place (lower; strict upper left
// untouched) and write the inverse of the factor (full n x n, upper part zero, ld = NBD)
// to Winv.  One CTA.  info: first failure wins (device-wide, stream ordered).
__global__ void __launch_bounds__(DIAG_THREADS, 1)
potrf_diag_kernel(int n, double* __restrict__ A, int lda, double* __restrict__ Winv, int* d_info, int info_base) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_info;
    const DiagSmem m = diag_smem(smem_raw);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int n32 = (n + SB - 1) / SB * SB;
    diag_load(m.S, A, lda, n, n32);
    __syncthreads();
    const int info = potrf_block_smem(m, n32, &s_info);
    if (info != 0 && tid == 0 && d_info) atomicCAS(d_info, 0, info_base + info);
    // store L (lower triangle only)
    for (int idx = tid; idx < n * n; idx += nt) {
        const int j = idx / n, i = idx - j * n;
        if (i >= j) A[size_t(j) * lda + i] = m.S[j * DPITCH + i];
    }
    trtri_assemble_smem(m, n32);
    trtri_store(m, n, Winv);
}

// Invert the nblk diagonal blocks (NBD x NBD, last one possibly smaller) of the lower
// triangular b x b matrix L: block q -> Winv + q*NBD*NBD (ld = NBD, upper part zero).
// One CTA per block.  Used by the stateless TRSM tile op, where only L arrives (W2:273-323).
__global__ void __launch_bounds__(DIAG_THREADS, 1)
trtri_diag_kernel(int b, const double* __restrict__ L, int ldl, double* __restrict__ Winv) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DiagSmem m = diag_smem(smem_raw);
    const int q = blockIdx.x;
    const int o = q * NBD;
    const int n = min(NBD, b - o);
    const int n32 = (n + SB - 1) / SB * SB;
    diag_load(m.S, L + size_t(o) * ldl + o, ldl, n, n32);
    __syncthreads();
    trtri_subblocks_smem(m, n32);
    trtri_assemble_smem(m, n32);
    trtri_store(m, n, Winv + size_t(q) * NBD * NBD);
}

}  // namespace chol
