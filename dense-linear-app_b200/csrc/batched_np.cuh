// batched_np.cuh — batched small Cholesky (configs[4]: 10 000 x 256), left-looking, WITHOUT a producer warp.
//
// Same algorithm and the same arithmetic (bit for bit) as potrf_batched_ll_kernel (batched.cuh); what changes is who
// feeds the slab ring.  In the plain kernel a fifth warp only issues TMA copies while holding a full warp's
// registers, which forces the 96-register cap (4 CTAs x 160 threads) and ~300 bytes of spills per thread: ncu counts
// 44 M local-memory instructions, 17 % of the kernel's L2 sectors.  Here a CTA is the four consumer warps alone: the
// LAST warp to finish with a slab (an atomic counter per stage in shared memory) re-arms the stage's mbarrier and
// issues the bulk copies of the slab STAGES ahead.  Four CTAs of 128 threads per SM may use 128 registers each.
// (Five such CTAs per SM at 96 registers were measured too: 4.76 against 4.73 ms — the larger working set overflows
// the L2: hit rate 57 -> 48.5 %, DRAM reads 7.3 -> 9.6 GB.)
#pragma once
#include "batched.cuh"

namespace chol {

constexpr int NP_THREADS = BL_CONSUMERS * 32;

template <int STAGES>
struct BatchedNP {
    static constexpr int SLAB = BLK * BL_PITCH;
    // slabs | L_jj | 1/l_cc | potrf32 column buffers | full barriers | release counters
    static constexpr size_t SMEM = size_t(STAGES * SLAB + BLW * BL_LP + BLW + 2 * SB) * 8 + STAGES * 8 + STAGES * 4 + 16;
};

template <int STAGES, int MIN_CTAS>
__global__ void __launch_bounds__(NP_THREADS, MIN_CTAS)
potrf_batched_np_kernel(int n, double* __restrict__ Abase, int lda, long long stride, int* __restrict__ d_info) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    constexpr int SLAB = BatchedNP<STAGES>::SLAB;
    double* slabs = reinterpret_cast<double*>(smem_dyn);
    double* Lcm = slabs + STAGES * SLAB;           // L_jj, column-major, pitch BL_LP
    double* invd = Lcm + BLW * BL_LP;              // 1 / l_cc
    double* bcast = invd + BLW;                    // potrf32_regs' column buffers (warp 0)
    uint64_t* bars = reinterpret_cast<uint64_t*>(bcast + 2 * SB);   // `full` barrier of every stage
    int* done = reinterpret_cast<int*>(bars + STAGES);               // consumer warps finished with the stage
    __shared__ int s_info;

    double* A = Abase + size_t(blockIdx.x) * size_t(stride);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
    const int g = lane >> 2, t = lane & 3;

    if (tid == 0) {
        s_info = 0;
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            done[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int it = 0;                                     // slabs streamed so far (same count in every warp)
    const int nb = n / BLW;
    for (int j = 0; j < nb; ++j) {
        const int r0 = j * BLW;
        const int R = n - r0;
        const int nslab = r0 / BLK;                 // K = 32 j columns of L to the left
        if (warp == BL_CONSUMERS - 1) {
            // warm L2 with the part of the matrix the next steps read from HBM: block column j + 1
            // (and, at the start, block column 0), so the accumulator / potrf / trsm loads hit L2
            if (j == 0) bulk_prefetch_l2(A + size_t(lane) * lda + lane / 2 * 2, uint32_t(n - lane / 2 * 2) * 8u);
            const int cn = r0 + BLW + lane;
            if (cn < n) bulk_prefetch_l2(A + size_t(cn) * lda + r0 + BLW, uint32_t(n - r0 - BLW) * 8u);
        }
        // one update pass: rows [row_first, row_first + prow) of block column j by the warps [w_first, 4).  Pass 0 is
        // run by every warp; pass 1 (rows beyond the first 128) by warps 1..3 — warp 0 factors the diagonal block
        // meanwhile.  A CTA barrier separates consecutive passes, so every stage is free when a pass starts.
        auto run_pass = [&](const int pass) {
            const int prow = pass == 0 ? min(R, BL_ROWS) : R - BL_ROWS;   // rows of C in this pass
            const int row_first = pass == 0 ? r0 : r0 + BL_ROWS;
            const int w_first = pass == 0 ? 0 : 1;
            const int w_cnt = BL_CONSUMERS - w_first;
            // slab = [32 B-rows | the pass' own rows] x 8 columns (pass 0: the B rows are its first 32 own rows)
            const uint32_t tx = uint32_t(BLK) * uint32_t(pass == 0 ? prow : BLW + prow) * 8u;
            auto fill = [&](const int s, const int st) {     // lanes 0..15 of the calling warp; lane 0 arms the barrier
                const uint32_t full = smem_u32(&bars[st]);
                if (lane == 0) mbar_expect_tx(full, tx);
                __syncwarp();
                const int col = lane & 7, part = lane >> 3;
                double* dst = slabs + st * SLAB + col * BL_PITCH;
                const double* src = A + size_t(s * BLK + col) * lda;
                if (pass == 0) {
                    if (part == 0) bulk_g2s(smem_u32(dst), src + r0, uint32_t(prow) * 8u, full);
                } else if (part == 0) {
                    bulk_g2s(smem_u32(dst), src + r0, uint32_t(BLW) * 8u, full);
                } else if (part == 1) {
                    bulk_g2s(smem_u32(dst + BLW), src + row_first, uint32_t(prow) * 8u, full);
                }
            };
            if (warp < w_first) {                   // (warp 0 during pass 1: keeps the slab counter in step)
                it += nslab;
                return;
            }
            // prologue: the first participating warp fills the ring
            if (warp == w_first)
                for (int s = 0; s < min(STAGES, nslab); ++s) fill(s, (it + s) % STAGES);
            const int blk = warp - w_first;
            const bool has = blk * BLW < prow;
            const int slab_row = pass == 0 ? blk * BLW : BLW + blk * BLW;
            // the accumulators start as the block of A itself and the products are subtracted by negating the B
            // fragments
            double acc[2][2][2][2][2];
            double* gC = A + size_t(r0) * lda + row_first + blk * BLW;
            if (has) {
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int e = 0; e < 2; ++e)
#pragma unroll
                            for (int np = 0; np < 2; ++np) {
                                const double2 v = *reinterpret_cast<const double2*>(
                                    gC + size_t(r * 16 + 4 * t + 2 * e + np) * lda + q * 16 + 2 * g);
                                acc[q][r][0][np][e] = v.x;
                                acc[q][r][1][np][e] = v.y;
                            }
            } else {
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int e = 0; e < 2; ++e)
#pragma unroll
                            for (int np = 0; np < 2; ++np) acc[q][r][0][np][e] = acc[q][r][1][np][e] = 0.0;
            }
            for (int s = 0; s < nslab; ++s, ++it) {
                const int st = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(smem_u32(&bars[st]), ph);
                if (has) {
                    const double* sl = slabs + st * SLAB + t * BL_PITCH + 2 * g;
#pragma unroll
                    for (int kk = 0; kk < BLK; kk += 4) {
                        double2 a[2], b[2];
#pragma unroll
                        for (int q = 0; q < 2; ++q)
                            a[q] = *reinterpret_cast<const double2*>(sl + kk * BL_PITCH + slab_row + q * 16);
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            b[r] = *reinterpret_cast<const double2*>(sl + kk * BL_PITCH + r * 16);
                            b[r].x = -b[r].x;
                            b[r].y = -b[r].y;
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q)
#pragma unroll
                            for (int r = 0; r < 2; ++r) {
                                dmma884(acc[q][r][0][0][0], acc[q][r][0][0][1], a[q].x, b[r].x);
                                dmma884(acc[q][r][0][1][0], acc[q][r][0][1][1], a[q].x, b[r].y);
                                dmma884(acc[q][r][1][0][0], acc[q][r][1][0][1], a[q].y, b[r].x);
                                dmma884(acc[q][r][1][1][0], acc[q][r][1][1][1], a[q].y, b[r].y);
                            }
                    }
                }
                // release: the last of the w_cnt warps to get here owns the stage and refills it
                __syncwarp();
                int last = 0;
                if (lane == 0) {
                    last = atomicAdd(&done[st], 1) == w_cnt - 1;
                    if (last) atomicExch(&done[st], 0);
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last && s + STAGES < nslab) fill(s + STAGES, st);
            }
            if (has) {
                // 2 consecutive rows per access; the diagonal block keeps its strict upper triangle untouched
                const bool diag = pass == 0 && blk == 0;
                // (opaque copy of the base: keeps the 16 store addresses from being computed before the main loop
                // and carried through it in registers)
                double* gS = gC;
                asm volatile("" : "+l"(gS));
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int e = 0; e < 2; ++e)
#pragma unroll
                            for (int np = 0; np < 2; ++np) {
                                const int rl = q * 16 + 2 * g;
                                const int cl = r * 16 + 4 * t + 2 * e + np;
                                double* ptr = gS + size_t(cl) * lda + rl;
                                if (diag && rl < cl) {
                                    if (rl + 1 == cl) ptr[1] = acc[q][r][1][np][e];
                                    continue;
                                }
                                *reinterpret_cast<double2*>(ptr) = make_double2(acc[q][r][0][np][e], acc[q][r][1][np][e]);
                            }
            }
        };
        const bool two = nslab > 0 && R > BL_ROWS;
        if (nslab > 0) {
            run_pass(0);
            __syncthreads();    // the updated diagonal block (pass 0, block 0) is visible to warp 0
        }
        // ---- diagonal block: one warp, registers + shared-memory broadcast; the others finish the update meanwhile
        if (two) run_pass(1);
        if (warp == 0) {
            double a[SB];
            const double* gD = A + size_t(r0) * lda + r0 + lane;
#pragma unroll
            for (int c = 0; c < SB; ++c) a[c] = (lane >= c) ? gD[size_t(c) * lda] : 0.0;
            double inv;
            // the factor goes to the shared copy L_jj (for the solve below), then from there to global memory
            const int info = potrf32_regs(a, inv, Lcm + lane, BL_LP, bcast);
            if (info != 0 && lane == 0 && s_info == 0) s_info = r0 + info;
            __syncwarp();
            double* gDw = A + size_t(r0) * lda + r0 + lane;
#pragma unroll
            for (int c = 0; c < SB; ++c)
                if (lane >= c) gDw[size_t(c) * lda] = Lcm[c * BL_LP + lane];
            invd[lane] = inv;
        }
        __syncthreads();
        // ---- rows below: X L_jj^T = C by forward substitution, one thread per row
        for (int row = r0 + BLW + tid; row < n; row += NP_THREADS) {
            // compiler barrier: without it the (loop-invariant) shared-memory loads of L_jj are hoisted out of
            // this loop and parked in local memory
            asm volatile("" ::: "memory");
            double x[SB];
            double* gX = A + size_t(r0) * lda + row;
#pragma unroll
            for (int c = 0; c < SB; ++c) x[c] = gX[size_t(c) * lda];
#pragma unroll
            for (int c = 0; c < SB; ++c) {
                const double xc = x[c] * invd[c];
                x[c] = xc;
                const double* lc = Lcm + c * BL_LP;
                if (((c + 1) & 1) != 0 && c + 1 < SB) x[c + 1] = fma(-xc, lc[c + 1], x[c + 1]);
#pragma unroll
                for (int jj = (c + 2) & ~1; jj < SB; jj += 2) {
                    const double2 l2 = *reinterpret_cast<const double2*>(lc + jj);
                    x[jj] = fma(-xc, l2.x, x[jj]);
                    x[jj + 1] = fma(-xc, l2.y, x[jj + 1]);
                }
            }
#pragma unroll
            for (int c = 0; c < SB; ++c) gX[size_t(c) * lda] = x[c];
        }
        // the next block column's TMA reads (async proxy) must see these generic-proxy writes
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncthreads();
    }
    if (tid == 0) d_info[blockIdx.x] = s_info;
}

}  // namespace chol
