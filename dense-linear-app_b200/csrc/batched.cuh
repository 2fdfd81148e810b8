// batched.cuh — many independent small Cholesky factorizations (the ArmoniK many-task
// workload: one POTRF task per small tile, C1:139-141 / W2:179-268, here one CTA per matrix).
#pragma once
#include <cuda_runtime.h>
#include "gemm_dmma.cuh"
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

// ---------------------------------------------------------------------------------------------------
// n % 32 == 0, 32 <= n <= 256: LEFT-LOOKING blocked kernel on the DMMA pipe, several matrices in
// flight per SM (configs[4]: 10 000 x 256).  One CTA (4 consumer warps + 1 producer warp) per
// matrix; for every 32-wide block column j:
//   1. update   C(R x 32) = A(r0:, j) - L(r0:, 0:K) L(r0:r0+32, 0:K)^T,  K = 32 j, on m8n8k4 DMMAs:
//      the K dimension streams through a ring of shared-memory slabs (8 columns of L each) filled by
//      the producer warp with TMA bulk copies out of this matrix's own, L2-hot, earlier columns; a
//      consumer warp owns one 32 x 32 block of C in registers.  Two passes of at most 128 rows keep
//      the accumulators at 32 doubles per thread (-> 4 CTAs per SM); the B operand (rows r0..r0+31)
//      is part of the first pass' slab and is re-staged for the second;
//   2. potrf32  the diagonal block, one warp, a row per lane in registers, warp shuffles;
//   3. trsm     the rows below it, one thread per row by forward substitution in axpy form against
//      the shared-memory copy of L_jj (more accurate than multiplying by an inverse, and it needs
//      none), written straight back to global memory.
// The matrix never has to fit in shared memory; HBM sees each lower-triangle element once in and
// once out (8 n (n+1) bytes per matrix = the algorithmic minimum), the re-reads of L hit L2.
//
// Measured on B200 with clock64() (one CTA of the four on an SM, 10 000 x 256, cycles per matrix): update
// 230k, potrf32 + second pass 190k, substitution 100k; the FP64/DMMA pipe is busy 36 % of the time, the rest
// is CTAs waiting at their own phase boundaries.  A variant with LOOKAHEAD inside the CTA (warps 1-3 already
// update block column j+1 by the columns < j while warp 0 factors block (j, j); every block of C visited
// twice, bit-identical results) was built and measured: faster for a single matrix (n = 224: 135 against
// 145 us) but slower at the full batch (5.14 against 4.73 ms) — the second visit re-streams the B rows and
// re-reads/re-writes C, and under four co-resident CTAs that extra L2 traffic costs more than the overlap
// gains.  Dropped.  What did help is dropping the PRODUCER WARP (batched_np.cuh, the default): the last consumer warp
// to release a slab refills it, a CTA is 128 threads, and four of them per SM may use 128 registers instead of 96 —
// the 300 bytes of spills per thread (44 M local-memory instructions, 17 % of the kernel's L2 sectors) shrink to
// 80: 4.43 against 4.73 ms, bit-identical.  FIVE such CTAs per SM at 96 registers: 4.76 ms — with 740 instead of
// 592 matrices in flight the working set (about 290 KB of touched lines per matrix) overflows the 126 MB L2 further
// (hit rate 57 -> 48.5 %, DRAM reads 7.3 -> 9.6 GB; minimum 2.6 GB).  Three per SM starve the pipes (5.56 ms).
// Getting to the 1.5 ms bound needs the matrix resident on chip while the serial chain of one matrix (8 x potrf32 at
// 5.6 us + 7 substitutions) is overlapped with the updates of another inside the same CTA.
constexpr int BLW = 32;                 // block-column width
constexpr int BLK = 8;                  // slab depth
constexpr int BL_ROWS = 128;            // rows per pass
constexpr int BL_PITCH = BL_ROWS + 4;   // doubles; same conflict-free LDS.128 pattern as gemm_dmma.cuh
constexpr int BL_CONSUMERS = 4;
constexpr int BL_THREADS = (BL_CONSUMERS + 1) * 32;
constexpr int BL_LP = 34;               // pitch of the column-major shared copy of L_jj (even: double2 loads)
constexpr int BATCHED_LL_MAX_N = 256;

template <int STAGES>
struct BatchedLL {
    static constexpr int SLAB = BLK * BL_PITCH;
    static constexpr size_t SMEM = size_t(STAGES * SLAB + BLW * BL_LP + BLW + 2 * SB) * 8 + 2 * STAGES * 8 + 16;
};

// In-register Cholesky of a 32x32 block, lane = row (a[j] = element (lane, j), upper part ignored).
// Returns the 1-based index of the first non-positive (or NaN) pivot, 0 if none; inv_out = 1 / l_cc of
// this lane's own diagonal element.
//
// The 32 pivots are one dependent chain, and on B200 a dependent FP64 operation issues ~35 cycles after
// its producer (measured with clock64(): the first version of this routine — pivot shuffle, rsqrt seed +
// two Newton steps, scaling, column shuffle, update — cost 370 cycles per pivot, 11.8k per block, and a
// variant that broadcast the column through shared memory instead of shuffles 410).  So the chain is kept
// to five FP64 operations per pivot:
//   * 1/sqrt(d): the hardware's 23-bit seed r0 (MUFU.RSQ64H) and ONE third-order correction
//     e = 1/2 - (d/2) r0^2,  1/sqrt(d) = r0 (1 + e + 3/2 e^2 + O(e^3)),  |e| <= 2^-22 so the cubic term is
//     below 2^-64 (1.2 units of 2^-53 in all, checked against 200-bit arithmetic);
//   * the scaled column is formed directly as l = x0 + (x0 e)(1 + 3/2 e), x0 = a r0 — the separate
//     multiplication by the refined reciprocal is gone (2.2 units of 2^-53);
//   * every lane keeps its OWN diagonal element in `diag`, updated with its own l (no shuffle), so the next
//     pivot is one shuffle behind the scaling instead of shuffle + fma + shuffle.
// sqrt(d) for the diagonal and the refined reciprocal for inv_out are computed off the chain.
//
// The loop over the pivots is a REAL loop (four phases of eight steps, `#pragma unroll 1`): fully unrolled,
// the compiler sinks every update of column j down to just before pivot j — the right-looking sweep turns
// into a left-looking one whose column update is a serial chain of j dependent FMAs (SASS of the unrolled
// form; 18.7k cycles per block).  To index registers statically inside a rolled loop the row is kept
// ROTATING: w[0] is always the current column, and the update writes column j into slot j-1.  The results go
// to `out` (column c of the factor at out[c * ldo], this lane's row) as they are produced; out may be
// shared or global memory.  `bc` = 64 doubles of shared memory private to the calling warp.  The next
// pivot's broadcast and rsqrt seed are issued before the row update.
template <int LIVE>
__device__ __forceinline__ void potrf32_phase(double (&w)[SB], double& diag, double& d, double& r0, int& info,
                                              double& inv_out, const int c0, double* __restrict__ out, const int ldo, double* bc) {
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int cc = 0; cc < 8; ++cc) {
        const int c = c0 + cc;
        if (!(d > 0.0) && info == 0) info = c + 1;
        const double e = fma(-(0.5 * d) * r0, r0, 0.5);
        const double v = fma(1.5, e, 1.0);
        const double x0 = w[0] * r0;
        const double l = fma(x0 * e, v, x0);              // on lane c: d / sqrt(d), unused by the others
        diag = fma(-l, l, diag);                          // lanes > c; lane c's own pivot is consumed
        const double dn = __shfl_sync(0xffffffffu, diag, (c + 1) & 31);
        double rn;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rn) : "d"(dn));
        {
            // off the chain: the refined reciprocal, and sqrt(d) to < 1 ulp for the diagonal element
            const double inv = fma(r0 * e, v, r0);
            const double piv = d * inv;
            const double root = fma(fma(-piv, piv, d), 0.5 * inv, piv);
            if (lane >= c) out[size_t(c) * ldo] = (lane == c) ? root : l;
            if (lane == c) inv_out = inv;
        }
        // row update, rotating: new w[j-1] = old w[j] - l * l_{c+j}.  The column reaches the other lanes through
        // shared memory (one STS, broadcast LDS): a single warp gets one SHFL through about every 11 cycles
        // (measured: 370 cycles per pivot with 2 x (31 - c) shuffles, whatever the length of the FP64 chain),
        // so shuffles carry only the pivot.  Two buffers: a lane may still be reading step c-1's column when
        // another one, past the __syncwarp of step c, writes step c+1's.
        double* col = bc + (c & 1) * SB;
        col[lane] = l;
        __syncwarp();
#pragma unroll
        for (int j = 1; j < LIVE; ++j) w[j - 1] = fma(-l, col[(c + j) & 31], w[j]);
        d = dn;
        r0 = rn;
    }
}

// a[j] = element (lane, j) on entry (upper part ignored); the factor is written to out / ldo (lower part only:
// lane >= column); bc = 64 doubles of shared scratch owned by this warp.  Returns info (first bad pivot, 1-based, 0 = none); inv_out = 1 / l_cc of this lane's own row.
__device__ __forceinline__ int potrf32_regs(double (&a)[SB], double& inv_out, double* __restrict__ out, const int ldo,
                                            double* bc) {
    const int lane = threadIdx.x & 31;
    int info = 0;
    inv_out = 0.0;
    double diag = 0.0;
#pragma unroll
    for (int c = 0; c < SB; ++c)
        if (lane == c) diag = a[c];
    double d = __shfl_sync(0xffffffffu, diag, 0);
    double r0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
    // columns beyond the block (wrapped lanes of the last steps' shuffles) only ever reach dead slots
    potrf32_phase<32>(a, diag, d, r0, info, inv_out, 0, out, ldo, bc);
    potrf32_phase<24>(a, diag, d, r0, info, inv_out, 8, out, ldo, bc);
    potrf32_phase<16>(a, diag, d, r0, info, inv_out, 16, out, ldo, bc);
    potrf32_phase<8>(a, diag, d, r0, info, inv_out, 24, out, ldo, bc);
    return info;
}

#ifdef CHOL_DIAG_CLOCKS   // debug builds only (tools/): phase timestamps of one CTA of the last batched launch
__device__ long long g_bat_clk[64];
#define BL_CLK(i) do { if (threadIdx.x == 0 && blockIdx.x == gridDim.x / 2) g_bat_clk[i] = clock64(); } while (0)
#else
#define BL_CLK(i) do { } while (0)
#endif

template <int STAGES, int MIN_CTAS>
__global__ void __launch_bounds__(BL_THREADS, MIN_CTAS)
potrf_batched_ll_kernel(int n, double* __restrict__ Abase, int lda, long long stride, int* __restrict__ d_info) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    constexpr int SLAB = BatchedLL<STAGES>::SLAB;
    double* slabs = reinterpret_cast<double*>(smem_dyn);
    double* Lcm = slabs + STAGES * SLAB;           // L_jj, column-major, pitch BL_LP
    double* invd = Lcm + BLW * BL_LP;              // 1 / l_cc
    double* bcast = invd + BLW;                    // potrf32_regs' column buffers (warp 0)
    uint64_t* bars = reinterpret_cast<uint64_t*>(bcast + 2 * SB);   // [0,STAGES) full, [STAGES,2*STAGES) empty
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
            mbar_init(smem_u32(&bars[STAGES + s]), BL_CONSUMERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int it = 0;                                     // slabs streamed so far (same count in every warp)
    const int nb = n / BLW;
    BL_CLK(0);
    for (int j = 0; j < nb; ++j) {
        const int r0 = j * BLW;
        const int R = n - r0;
        const int nslab = r0 / BLK;                 // K = 32 j columns of L to the left
        if (warp == BL_CONSUMERS) {
            // warm L2 with the part of the matrix the next steps read from HBM: block column j + 1
            // (and, at the start, block column 0), so the epilogue / potrf / trsm loads hit L2
            if (j == 0) bulk_prefetch_l2(A + size_t(lane) * lda + lane / 2 * 2, uint32_t(n - lane / 2 * 2) * 8u);
            const int cn = r0 + BLW + lane;
            if (cn < n) bulk_prefetch_l2(A + size_t(cn) * lda + r0 + BLW, uint32_t(n - r0 - BLW) * 8u);
        }
        // one update pass: rows [row_first, row_first + prow) of block column j.  Pass 0 is run by
        // every warp; pass 1 (rows beyond the first 128) by warps 1..3 and the producer only — warp 0
        // factors the diagonal block meanwhile, and warp 1 releases the slabs on its behalf.
        auto run_pass = [&](const int pass) {
            {
                const int prow = pass == 0 ? min(R, BL_ROWS) : R - BL_ROWS;   // rows of C in this pass
                const int row_first = pass == 0 ? r0 : r0 + BL_ROWS;
                if (warp == BL_CONSUMERS) {
                    // ---------------- producer: slab = [32 B-rows | the pass' own rows] x 8 columns
                    const int col = lane & 7;
                    const int part = lane >> 3;     // 0: B rows (pass 1) / all rows (pass 0); 1: own rows of pass 1
                    const uint32_t tx = uint32_t(BLK) * uint32_t(pass == 0 ? prow : BLW + prow) * 8u;
                    for (int s = 0; s < nslab; ++s, ++it) {
                        const int st = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        mbar_wait(smem_u32(&bars[STAGES + st]), ph ^ 1);
                        const uint32_t full = smem_u32(&bars[st]);
                        if (lane == 0) mbar_expect_tx(full, tx);
                        __syncwarp();
                        double* dst = slabs + st * SLAB + col * BL_PITCH;
                        const double* src = A + size_t(s * BLK + col) * lda;
                        if (pass == 0) {
                            if (part == 0) bulk_g2s(smem_u32(dst), src + r0, uint32_t(prow) * 8u, full);
                        } else if (part == 0) {
                            bulk_g2s(smem_u32(dst), src + r0, uint32_t(BLW) * 8u, full);
                        } else if (part == 1) {
                            bulk_g2s(smem_u32(dst + BLW), src + row_first, uint32_t(prow) * 8u, full);
                        }
                    }
                } else {
                    // ---------------- consumers: one 32 x 32 block of C per warp
                    // pass 0: warp w -> rows [32 w, 32 w + 32) of the slab; pass 1: warps 1..3 -> own rows
                    const int blk = pass == 0 ? warp : warp - 1;
                    const bool has = blk >= 0 && blk * BLW < prow;
                    const int slab_row = pass == 0 ? blk * BLW : BLW + blk * BLW;
                    // the accumulators start as the block of A itself (its HBM latency hides behind the
                    // first slab's) and the products are subtracted by negating the B fragments
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
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(smem_u32(&bars[STAGES + st]));
                            if (pass == 1 && warp == 1) mbar_arrive(smem_u32(&bars[STAGES + st]));   // for warp 0
                        }
                    }
                    if (has) {
                        // 2 consecutive rows per access; the diagonal block keeps its strict upper
                        // triangle untouched
                        const bool diag = pass == 0 && blk == 0;
                        // (opaque copy of the base: keeps the 16 store addresses from being computed
                        // before the main loop and carried through it in registers -> spills)
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
                                        *reinterpret_cast<double2*>(ptr) =
                                            make_double2(acc[q][r][0][np][e], acc[q][r][1][np][e]);
                                    }
                    }
                }
            }
        };
        const bool two = nslab > 0 && R > BL_ROWS;
        if (nslab > 0) {
            run_pass(0);
            __syncthreads();    // the updated diagonal block (pass 0, block 0) is visible to warp 0
        }
        BL_CLK(1 + 3 * j);
        // ---- diagonal block: one warp, registers + shuffles; the others finish the update meanwhile
        if (warp != 0) {
            if (two) run_pass(1);
        } else {
            if (two) it += nslab;
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
        BL_CLK(2 + 3 * j);
        // ---- rows below: X L_jj^T = C by forward substitution, one thread per row
        for (int row = r0 + BLW + tid; row < n; row += BL_THREADS) {
            // compiler barrier: without it the (loop-invariant) shared-memory loads of L_jj are hoisted
            // out of this loop and parked in local memory (4.6 KB of spills)
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
        BL_CLK(3 + 3 * j);
    }
    if (tid == 0) d_info[blockIdx.x] = s_info;
}

}  // namespace chol
