// panel2.cuh — second-generation panel kernels for tiles whose size is a multiple of 32 (every
// BASELINE configuration): the diagonal-block Cholesky and the triangular-solve leaf.
//
// Round 1's kernels (panel.cuh) built the full inverse of every 128x128 diagonal block and multiplied
// by it with a 198 KB / 288-thread GEMM CTA.  Measured on B200 (profiles/r02_trace_*): a POTRF tile
// takes 0.95 ms alone and 2.0-2.9 ms while the trailing update runs, a panel TRSM 0.45 / 3.3 ms —
// the whole-SM CTAs of the chain wait for BOTH resident update CTAs of an SM to retire at every one
// of the 16 dependent launches.  Here:
//   * potrf_diag32_kernel: the 128x128 block lives in shared memory as packed 32x32 blocks (92 KB,
//     256 threads: fits beside a resident update CTA); one warp factors a 32x32 block in registers
//     (rsqrt pivots, batched shuffles), the rows below are solved by forward SUBSTITUTION (one
//     thread per row) while another warp inverts the 32x32 block for later use, the trailing
//     blocks are updated on the DMMA pipe.  Only the four 32x32 inverses D_J are produced — no
//     128x128 inverse assembly;
//   * trsm_leaf32_kernel: X = A L_jj^{-T} for a 128-column block by block substitution with the
//     D_J (X_J = A_J D_J^T; A_I -= X_J L_IJ^T), every warp on its own 16 rows held in DMMA
//     accumulators from load to store — no barrier wider than a warp, 20 KB of shared memory, half
//     the flops of the multiply by a full inverse, and only 32x32 inverses enter the arithmetic.
#pragma once
#include <cuda_runtime.h>
#include "batched.cuh"
#include "gemm_dmma.cuh"
#include "panel.cuh"

namespace chol {

constexpr int D2_THREADS = 256;
constexpr int D2_BP = 36;                         // pitch of a packed 32x32 block: == 4 (mod 16) doubles
constexpr int D2_BLK = SB * D2_BP;                // doubles per block
constexpr int D2_NBLK = 10;                       // lower blocks of a 4x4 block matrix
constexpr size_t D2_SMEM = size_t(D2_NBLK * D2_BLK + NBD + 2 * SB) * 8 + 16;

#ifdef CHOL_DIAG_CLOCKS   // debug builds only (tools/): phase timestamps of the last diag32 launch
__device__ long long g_diag_clk[64];
#define D2_CLK(i) do { if (threadIdx.x == 0) g_diag_clk[i] = clock64(); } while (0)
#else
#define D2_CLK(i) do { } while (0)
#endif

__device__ __forceinline__ int d2_blk(int I, int J) { return I * (I + 1) / 2 + J; }   // I >= J

// Factor the n x n (n % 32 == 0, n <= 128) diagonal block at A in place (lower; strict upper
// untouched) and write D_J = inv(L_JJ) of its 32x32 diagonal blocks into Winv (ld NBD, at their
// natural position; lower triangle, upper part zero).
__global__ void __launch_bounds__(D2_THREADS, 2)
potrf_diag32_kernel(int n, double* __restrict__ A, int lda, double* __restrict__ Winv, int* d_info, int info_base) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* S = reinterpret_cast<double*>(smem_raw);
    double* invd = S + D2_NBLK * D2_BLK;
    double* bcast = invd + NBD;                     // potrf32_regs' column buffers (warp 0)
    __shared__ int s_info;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = lane >> 2, t = lane & 3;
    const int nb = n / SB;
    if (tid == 0) s_info = 0;
    D2_CLK(0);
    // ---- load the lower blocks with the TMA engine: one 256-byte bulk copy per block column (a column of a
    // packed block is contiguous in both memories), all in flight at once, completion on one mbarrier.
    // (The per-thread loads this replaces took 4.4k cycles, two serialized L2 round trips.)  The strict upper
    // triangle of the diagonal blocks comes along and is ignored.
    {
        const uint32_t bar = smem_u32(&s_bar);
        const int ncols = nb * (nb + 1) / 2 * SB;
        pdl_sync();
        if (tid == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(bar, uint32_t(ncols) * uint32_t(SB * 8));
        }
        __syncthreads();
        for (int idx = tid; idx < ncols; idx += D2_THREADS) {
            const int bi = idx >> 5, c = idx & 31;
            int I = 0;
            while ((I + 1) * (I + 2) / 2 <= bi) ++I;
            const int J = bi - I * (I + 1) / 2;
            bulk_g2s(smem_u32(S + bi * D2_BLK + c * D2_BP), A + size_t(J * SB + c) * lda + I * SB, SB * 8, bar);
        }
        mbar_wait(bar, 0);
    }
    __syncthreads();
    D2_CLK(1);
    for (int J = 0; J < nb; ++J) {
        double* LJJ = S + d2_blk(J, J) * D2_BLK;
        if (warp == 0) {
            double a[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c) a[c] = (lane >= c) ? LJJ[c * D2_BP + lane] : 0.0;
            double inv;
            // (only the lower part of L_JJ is written, and nothing reads the strict upper part)
            const int info = potrf32_regs(a, inv, LJJ + lane, D2_BP, bcast);
            if (info != 0 && lane == 0 && s_info == 0) s_info = J * SB + info;
            invd[J * SB + lane] = inv;
        }
        __syncthreads();
        D2_CLK(2 + 3 * J);
        const int R = (nb - 1 - J) * SB;           // rows below: 0, 32, 64 or 96
        if (tid < R) {
            // ---- forward substitution, one thread per row of the blocks (I, J), I > J
            const int I = J + 1 + (tid >> 5);
            double* X = S + d2_blk(I, J) * D2_BLK + lane;
            double x[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c) x[c] = X[c * D2_BP];
#pragma unroll
            for (int c = 0; c < SB; ++c) {
                const double xc = x[c] * invd[J * SB + c];
                x[c] = xc;
                const double* lc = LJJ + c * D2_BP;
                if (((c + 1) & 1) != 0 && c + 1 < SB) x[c + 1] = fma(-xc, lc[c + 1], x[c + 1]);
#pragma unroll
                for (int jj = (c + 2) & ~1; jj < SB; jj += 2) {
                    const double2 l2 = *reinterpret_cast<const double2*>(lc + jj);
                    x[jj] = fma(-xc, l2.x, x[jj]);
                    x[jj + 1] = fma(-xc, l2.y, x[jj + 1]);
                }
            }
#pragma unroll
            for (int c = 0; c < SB; ++c) X[c * D2_BP] = x[c];
        } else if (warp == 7) {
            // ---- meanwhile: D_J = inv(L_JJ), lane j owns column j (forward substitution, axpy form),
            // written straight to global memory
            // (1 / l_kk by division from the stored factor, exactly as the stateless TRSM op derives it
            // from L alone in trtri_diag_kernel: both forms of the solve then agree bit for bit)
            const double myinv = 1.0 / LJJ[lane * D2_BP + lane];
            double acc[SB];
#pragma unroll
            for (int i = 0; i < SB; ++i) acc[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < SB; ++k) {
                const double w = acc[k] * __shfl_sync(0xffffffffu, myinv, k);
                acc[k] = w;
#pragma unroll
                for (int i = k + 1; i < SB; ++i) acc[i] = fma(-LJJ[k * D2_BP + i], w, acc[i]);
            }
            double* Wd = Winv + size_t(J * SB + lane) * NBD + J * SB;
#pragma unroll
            for (int i = 0; i < SB; ++i) Wd[i] = acc[i];
        }
        if (R == 0) break;
        __syncthreads();
        D2_CLK(3 + 3 * J);
        // ---- trailing update on the DMMA pipe: block (I, K) -= X_I X_K^T, J < K <= I, cut into 16 x 16
        // quarters dealt round-robin to the 8 warps.  (One warp per 32 x 32 block left the DMMA pipe of two
        // of the four sub-partitions with two blocks and the others with one or none: 8.5k / 6.2k / 5.5k
        // cycles for 6 / 3 / 1 blocks; a quarter is 32 DMMAs = 512 pipe cycles.)  Same products, same
        // order of accumulation per element as the whole-block form.
        {
            const int m = nb - 1 - J;              // 1..3 -> m (m + 1) / 2 <= 6 blocks
            const int nq = 2 * m * (m + 1);
            for (int qt = warp; qt < nq; qt += D2_THREADS / 32) {
                const int task = qt >> 2, q = (qt >> 1) & 1, r = qt & 1;
                int ti = 0;
                while ((ti + 1) * (ti + 2) / 2 <= task) ++ti;
                const int I = J + 1 + ti, K = J + 1 + (task - ti * (ti + 1) / 2);
                if (I == K && q == 0 && r == 1) continue;      // strict upper quarter of a diagonal block: never read
                const double* XI = S + d2_blk(I, J) * D2_BLK + t * D2_BP + q * 16 + 2 * g;
                const double* XK = S + d2_blk(K, J) * D2_BLK + t * D2_BP + r * 16 + 2 * g;
                double* C = S + d2_blk(I, K) * D2_BLK + (r * 16 + 4 * t) * D2_BP + q * 16 + 2 * g;
                double acc[2][2][2];               // [mp][np][e]
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                    for (int np = 0; np < 2; ++np) {
                        const double2 v = *reinterpret_cast<const double2*>(C + (2 * e + np) * D2_BP);
                        acc[0][np][e] = v.x;
                        acc[1][np][e] = v.y;
                    }
#pragma unroll
                for (int kk = 0; kk < SB; kk += 4) {
                    const double2 a = *reinterpret_cast<const double2*>(XI + kk * D2_BP);
                    double2 b = *reinterpret_cast<const double2*>(XK + kk * D2_BP);
                    b.x = -b.x;
                    b.y = -b.y;
                    dmma884(acc[0][0][0], acc[0][0][1], a.x, b.x);
                    dmma884(acc[0][1][0], acc[0][1][1], a.x, b.y);
                    dmma884(acc[1][0][0], acc[1][0][1], a.y, b.x);
                    dmma884(acc[1][1][0], acc[1][1][1], a.y, b.y);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                    for (int np = 0; np < 2; ++np)
                        *reinterpret_cast<double2*>(C + (2 * e + np) * D2_BP) = make_double2(acc[0][np][e], acc[1][np][e]);
            }
        }
        __syncthreads();
        D2_CLK(4 + 3 * J);
    }
    __syncthreads();
    D2_CLK(14);
    if (s_info != 0 && tid == 0 && d_info) atomicCAS(d_info, 0, info_base + s_info);
    // ---- store L, lower triangle only, with bulk copies shared memory -> global: a whole column of an
    // off-diagonal block, rows c.. of column c of a diagonal block (from an even row: 16-byte alignment;
    // the diagonal element of an odd column goes out as a plain store).  The per-thread LDS + STG loop this
    // replaces took 8k cycles.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the blocks were written through the generic proxy
    __syncthreads();
    {
        const int ncols = nb * (nb + 1) / 2 * SB;
        for (int idx = tid; idx < ncols; idx += D2_THREADS) {
            const int bi = idx >> 5, c = idx & 31;
            int I = 0;
            while ((I + 1) * (I + 2) / 2 <= bi) ++I;
            const int J = bi - I * (I + 1) / 2;
            const double* src = S + bi * D2_BLK + c * D2_BP;
            double* dst = A + size_t(J * SB + c) * lda + I * SB;
            int r0 = 0;
            if (I == J) {
                r0 = (c + 1) & ~1;                                 // first even row >= c
                if (r0 != c) dst[c] = src[c];
            }
            if (r0 < SB) bulk_s2g(dst + r0, smem_u32(src + r0), uint32_t(SB - r0) * 8u);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    D2_CLK(15);
}

// ---------------------------------------------------------------------------------------------------
// X = A L_jj^{-T} in place for a 128-column (or narrower, multiple of 32) block of `m` rows.
// L points at the top-left of L_jj inside the factored tile (ld ldl); Dinv at this block's slot of the
// POTRF workspace (ld NBD; only the diagonal 32x32 blocks are read).
//
// The B operands (D_J and the blocks L_IJ, 8 KB each, shared by every warp of every CTA) are read
// straight from global memory through the read-only path as DMMA fragments: 16-byte loads, each
// warp-wide load covers four full 128-byte lines, and they stay in L1/L2.  (A first version staged
// them through a TMA + mbarrier ring; under a concurrent trailing update roughly one launch in a
// hundred returned a wrong 16- or 64-row strip — tools/gpu_trsm_race.py — and the version with plain
// loads never did, so the ring is gone; the blocks are too small to need it.)
constexpr int LF_ROWS = 64;                       // rows per CTA: 4 warps x 16 rows
constexpr int LF_THREADS = 128;
constexpr int LF_SP = 20;                         // staging pitch (16 rows + 4): == 4 (mod 16) doubles
constexpr size_t LF_SMEM = size_t(4 * SB * LF_SP) * 8;

struct LeafParams {
    double* const* tile_ptrs;   // panel mode: task t works on tile_ptrs[t] + off;  nullptr -> `single`
    double* single;
    long long off;              // doubles: column offset of this 128-block inside the tile
    int m, nbk, lda;            // rows, number of 32-column blocks (1..4), leading dimension of the tiles
    const double* L;
    int ldl;
    const double* Dinv;
    int ctas_per_task;
    // fused panel push (multi-GPU): task t also stores its result into the receive slots of up to
    // LF_MAX_PEERS ranks — peer_dst[t * npeer + q] is the address of this tile in peer q's mapped
    // slot buffer, 0 = that peer does not read this tile.  The transfer over NVLink then overlaps the
    // solve column block by column block instead of following it as a copy.
    const long long* peer_dst;
    int npeer;
};
constexpr int LF_MAX_PEERS = 7;

// one 16 x 32 x 32 product on the DMMA pipe: acc (+)= stg(16 x 32, A operand) * B^T, B (32 x 32) column-major
// at `B` with leading dimension ldb in global memory (read-only during the kernel); NEG subtracts.
template <bool NEG>
__device__ __forceinline__ void leaf_product(double (&acc)[2][2][2][2], const double* stg, const double* __restrict__ B,
                                             int ldb, int g, int t) {
    const double* Bl = B + size_t(t) * ldb + 2 * g;
#pragma unroll
    for (int kk = 0; kk < SB; kk += 4) {
        const double2 a = *reinterpret_cast<const double2*>(stg + (kk + t) * LF_SP + 2 * g);
        double2 b[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            b[r] = __ldg(reinterpret_cast<const double2*>(Bl + size_t(kk) * ldb + r * 16));
            if (NEG) {
                b[r].x = -b[r].x;
                b[r].y = -b[r].y;
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            dmma884(acc[r][0][0][0], acc[r][0][0][1], a.x, b[r].x);
            dmma884(acc[r][0][1][0], acc[r][0][1][1], a.x, b[r].y);
            dmma884(acc[r][1][0][0], acc[r][1][0][1], a.y, b[r].x);
            dmma884(acc[r][1][1][0], acc[r][1][1][1], a.y, b[r].y);
        }
    }
}

__global__ void __launch_bounds__(LF_THREADS, 2) trsm_leaf32_kernel(const __grid_constant__ LeafParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* stage_all = reinterpret_cast<double*>(smem_raw);                    // 4 warps x (32 cols x LF_SP)
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = lane >> 2, t = lane & 3;
    const int task = blockIdx.x / p.ctas_per_task;
    const int rb = blockIdx.x - task * p.ctas_per_task;
    double* Abase = (p.tile_ptrs ? p.tile_ptrs[task] : p.single) + p.off;
    const int nbk = p.nbk;
    pdl_sync();      // (the tile-pointer list is plan data)
    if ((reinterpret_cast<uintptr_t>(Abase) & 15u) != 0) {
        // a tile of the pointer list that is only 8-byte aligned (the host cannot see device pointer lists):
        // plain forward substitution, one thread per row — slow, correct, and no misaligned 16-byte access
        const int ncol = nbk * SB;
        for (int row = rb * LF_ROWS + tid; row < min(p.m, (rb + 1) * LF_ROWS); row += LF_THREADS) {
            double* a = Abase + row;
            for (int c = 0; c < ncol; ++c) {
                double sacc = a[size_t(c) * p.lda];
                for (int k = 0; k < c; ++k) sacc = fma(-a[size_t(k) * p.lda], p.L[size_t(k) * p.ldl + c], sacc);
                a[size_t(c) * p.lda] = sacc / p.L[size_t(c) * p.ldl + c];
            }
        }
        return;
    }
    // warp w owns rows row0 .. row0 + 15 of this CTA's strip, as DMMA accumulators, from load to store
    const int row0 = rb * LF_ROWS + warp * 16;
    const bool live = row0 + 2 * g < p.m;                  // m is even: the row pair is in or out together
    double* gA = Abase + row0 + 2 * g;
    double* stg = stage_all + warp * SB * LF_SP;
    double* gP[LF_MAX_PEERS];                               // the same rows of this tile in the peers' slots
#pragma unroll
    for (int q = 0; q < LF_MAX_PEERS; ++q) {
        long long d = 0;
        if (q < p.npeer) d = __ldg(p.peer_dst + size_t(task) * p.npeer + q);
        gP[q] = d ? reinterpret_cast<double*>(d) + p.off + row0 + 2 * g : nullptr;
    }
    double acc[4][2][2][2][2];                              // [block][r][mp][np][e]
#pragma unroll
    for (int I = 0; I < 4; ++I)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    double2 v = make_double2(0.0, 0.0);
                    if (live && I < nbk)
                        v = *reinterpret_cast<const double2*>(gA + size_t(I * SB + r * 16 + 4 * t + 2 * e + np) * p.lda);
                    acc[I][r][0][np][e] = v.x;
                    acc[I][r][1][np][e] = v.y;
                }
#pragma unroll
    for (int J = 0; J < 4; ++J) {
        if (J < nbk) {
            // A_J (accumulator layout) -> staging as an A operand: stg[k * LF_SP + row]
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                    for (int np = 0; np < 2; ++np)
                        *reinterpret_cast<double2*>(stg + (r * 16 + 4 * t + 2 * e + np) * LF_SP + 2 * g) =
                            make_double2(acc[J][r][0][np][e], acc[J][r][1][np][e]);
            __syncwarp();
            // X_J = A_J D_J^T
            double x[2][2][2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int mp = 0; mp < 2; ++mp)
#pragma unroll
                    for (int np = 0; np < 2; ++np) x[r][mp][np][0] = x[r][mp][np][1] = 0.0;
            leaf_product<false>(x, stg, p.Dinv + size_t(J * SB) * NBD + J * SB, NBD, g, t);
            __syncwarp();       // every lane is done reading A_J from the staging buffer
            // X_J: final -> global memory, and -> staging (A operand of the updates)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                    for (int np = 0; np < 2; ++np) {
                        const double2 v = make_double2(x[r][0][np][e], x[r][1][np][e]);
                        const int c = r * 16 + 4 * t + 2 * e + np;
                        *reinterpret_cast<double2*>(stg + c * LF_SP + 2 * g) = v;
                        if (live) {
                            *reinterpret_cast<double2*>(gA + size_t(J * SB + c) * p.lda) = v;
#pragma unroll
                            for (int q = 0; q < LF_MAX_PEERS; ++q)
                                if (gP[q]) *reinterpret_cast<double2*>(gP[q] + size_t(J * SB + c) * p.lda) = v;
                        }
                    }
            __syncwarp();
            // A_I -= X_J L_IJ^T for the blocks to the right
#pragma unroll
            for (int I = J + 1; I < 4; ++I)
                if (I < nbk) leaf_product<true>(acc[I], stg, p.L + size_t(J * SB) * p.ldl + I * SB, p.ldl, g, t);
            __syncwarp();       // every lane is done reading X_J before the staging buffer is rewritten
        }
    }
}

}  // namespace chol
