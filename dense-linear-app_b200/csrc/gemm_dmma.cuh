// gemm_dmma.cuh — grouped FP64 rank-K update on the sm_100a DMMA path.
//
//   C_t <- beta*C_t + alpha * A_t * B_t^T        for every task t of a task list
//
// This one kernel is the hot path of the tiled Cholesky: with alpha=-1, beta=1 it is the
// GEMM tile op (W2:511 CHAMELEON_dgemm_Tile(NoTrans,Trans,-1,Ai,Aj,1,C)), with the LOWER
// flag the SYRK tile op (W2:416 CHAMELEON_dsyrk_Tile(Lower,NoTrans,-1,A,1,C)), and a task
// list holding every (i,j) of one wave of the client loop (C1:307-329) is the fused
// per-panel trailing update.  With alpha=1, beta=0 and B = inverse of a diagonal block it
// is the multiply step of the blocked TRSM (W2:323).
//
// Design (B200):
//   * CTA tile 128 x BN, K sliced in slabs of 16; consumer warps (2 x BN/32) each own a 64x32
//     sub-tile held in registers as 32 m8n8k4 DMMA accumulators (64 doubles);
//   * 1 producer warp stages slabs with the TMA engine: one `cp.async.bulk` (UBLKCP) per
//     slab column (a column of a col-major tile is contiguous), completion counted on an
//     mbarrier (full/empty ring);
//   * two shapes (GemmCfg below): BN=64, 4 consumer warps, 4 stages (100 KB), TWO CTAs per SM for
//     the trailing update — one CTA's pipeline fill and C epilogue hide behind the other's main
//     loop (ncu: DMMA pipe 85.7 -> 92.8 % active); BN=128, 8 consumer warps, 6 stages (198 KB), one
//     CTA per SM for the in-place multiply steps of the blocked TRSM/POTRF;
//   * slab columns are placed at a pitch of BM+4 / BN+4 doubles: the fragment loads are
//     LDS.128 with lane -> (k = lane&3, rows 2*(lane>>2)..+1), which those pitches make
//     bank-conflict free (each quarter warp covers all 8 16-byte bank groups);
//   * one LDS.128 feeds two DMMAs (even/odd rows), so a k4 step is 6 LDS.128 : 32 DMMA;
//   * C is read-modify-written straight from the accumulators: each lane owns 2x4
//     patches (2 consecutive rows x 4 consecutive columns) -> 16-byte accesses, every
//     warp-wide access covers four full 128-byte lines; the loads of a 16-row group are issued
//     together, and the producer prefetches the C block into L2 while the main loop runs.
//
// DRAM traffic (ncu, B200): a 120-task update at b = 1024 reads 1.90 GB and writes 0.91 GB against 2.14 GB
// algorithmic (C once each way + 15 panel tiles once) = 1.31x; an 820-task one 13.65 + 6.68 GB against 14.1 GB.
// The excess is the B operand: with tasks in row-major (i, j) order a row of tasks walks over every panel tile,
// and a panel of 15-40 tiles (126-335 MB) does not survive in L2 between rows — shared read-only lines are
// replicated in both halves of the 126 MB L2, so ~63 MB are effective.  L2 eviction hints (evict_last on the
// operand copies, evict_first on the C accesses) were measured and change nothing (1.88 against 1.90 GB,
// 34.2 against 34.3 TFLOP/s).  At 0.4 TB/s this traffic is 6 % of HBM bandwidth and costs no time.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/chol_b200.h"

namespace chol {

constexpr int BM = 128, BK = 16;
constexpr int PITCH = BM + 4;                       // doubles per A slab column in smem

// Two shapes of the same kernel:
//   GemmWide   128x128 CTA tile, 8 consumer warps, 1 CTA/SM  - the in-place multiply steps of the
//              blocked TRSM/POTRF (they need the whole 128-column block in ONE CTA, see launch_gemm);
//   GemmPair   128x64 CTA tile, 4 consumer warps, 2 CTAs/SM  - the trailing update: while one CTA
//              reads/writes its C block (epilogue) or fills its pipeline, the co-resident CTA
//              keeps the DMMA pipe busy (one warp per SM sub-partition saturates it).
template <int BN_, int STAGES_, int MIN_CTAS_>
struct GemmCfg {
    static constexpr int BN = BN_;
    static constexpr int STAGES = STAGES_;
    static constexpr int MIN_CTAS = MIN_CTAS_;
    static constexpr int WARPS_N = BN_ / 32;
    static constexpr int CONSUMER_WARPS = 2 * WARPS_N;
    static constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
    static constexpr int PITCH_B = BN_ + 4;
    static constexpr int SLAB_A = BK * PITCH;
    static constexpr int SLAB_B = BK * PITCH_B;
    static constexpr int STAGE_DOUBLES = SLAB_A + SLAB_B;
    static constexpr size_t SMEM_BYTES = size_t(STAGES_) * STAGE_DOUBLES * 8 + 2 * STAGES_ * 8 + 16;
};
using GemmWide = GemmCfg<128, 6, 1>;
using GemmPair = GemmCfg<64, 4, 2>;

struct GemmParams {
    const chol_task_t* tasks;   // device array, or nullptr -> `tile_ptrs` / `one`
    // panel mode (blocked TRSM over a list of tiles): task t is C = tile[t] + c_off,
    // A = tile[t] + a_off, B = one.B, flags = 0 — no task list has to be materialised
    double* const* tile_ptrs;
    long long c_off, a_off;
    chol_task_t one;
    int ntasks;
    int m, n, k;
    int lda, ldb, ldc;
    int nbm, nbn;               // CTA blocks per task
    double alpha, beta;
};

// ---- PTX helpers -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// TMA engine 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
// Programmatic dependent launch (the latency-bound panel chain: 24 dependent launches per POTRF tile, 15 per panel
// TRSM).  A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may become resident while its
// predecessor in the stream is still running; it must call this before it touches anything the predecessor wrote.
// The trigger for ITS successor comes right after the wait, so at most one successor is resident early (a trigger
// at the very top would let the whole chain pile up on the SMs).  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// TMA engine 1-D bulk copy shared -> global (bulk async-group completion).
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// D(8x8) += A(8x4, row) * B(4x8, col), FP64 tensor core (SASS: DMMA.8x8x4).
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ chol_task_t load_task(const GemmParams& p, int task_id) {
    chol_task_t task;
    if (p.tasks) {
        // 32-byte task record, two 16-byte loads
        const int4* tp = reinterpret_cast<const int4*>(p.tasks + task_id);
        int4 lo = __ldg(tp), hi = __ldg(tp + 1);
        task.C = reinterpret_cast<double*>((uint64_t(uint32_t(lo.y)) << 32) | uint32_t(lo.x));
        task.A = reinterpret_cast<const double*>((uint64_t(uint32_t(lo.w)) << 32) | uint32_t(lo.z));
        task.B = reinterpret_cast<const double*>((uint64_t(uint32_t(hi.y)) << 32) | uint32_t(hi.x));
        task.flags = (int64_t(hi.w) << 32) | uint32_t(hi.z);
    } else if (p.tile_ptrs) {
        double* tile = p.tile_ptrs[task_id];
        task.C = tile + p.c_off;
        task.A = tile + p.a_off;
        task.B = p.one.B;
        task.flags = 0;
    } else {
        task = p.one;
    }
    return task;
}

// Slow path of one CTA tile for tasks whose pointers are not 16-byte aligned.  Inlined on purpose: the CTA
// returns right after it, so its few registers never overlap the accumulators' live range, while an
// out-of-line call gives the kernel a stack frame and ABI spills (ptxas: 96 B stack, 80 B spill stores).
__device__ __forceinline__ void gemm_block_scalar(const GemmParams& p, const chol_task_t& task, bool lower, int row0,
                                               int col0, int mv, int nv) {
    for (int idx = threadIdx.x; idx < mv * nv; idx += blockDim.x) {
        const int i = row0 + idx % mv, j = col0 + idx / mv;
        if (lower && i < j) continue;
        double s = 0.0;
        for (int l = 0; l < p.k; ++l) s = fma(task.A[size_t(l) * p.lda + i], task.B[size_t(l) * p.ldb + j], s);
        double* c = task.C + size_t(j) * p.ldc + i;
        double v = p.alpha * s;
        if (p.beta != 0.0) v += p.beta * *c;
        *c = v;
    }
}

// ---- the kernel --------------------------------------------------------------------
// One CTA per 128 x BN tile, scheduled by the hardware.  A persistent variant (grid = resident CTAs, each
// walking tiles grid-stride with the producer warp running ahead into the next tile) was measured on B200 and
// dropped: 31.4 / 32.5 TFLOP/s against 33.5 / 34.0 for this form at 120 / 820 tasks — the two CTAs of an SM
// fall into step and reach their epilogues together, and the per-tile task fetch sits on the consumers' path,
// while fresh CTAs handed out by the hardware stay staggered.
#ifndef CHOL_GEMM_MAXNREG
#define CHOL_GEMM_MAXNREG 0
#endif
template <class Cfg>
__global__ void
#if CHOL_GEMM_MAXNREG > 0
__maxnreg__(CHOL_GEMM_MAXNREG)
#else
__launch_bounds__(Cfg::THREADS, Cfg::MIN_CTAS)
#endif
gemm_nt_dmma_kernel(const __grid_constant__ GemmParams p) {
    constexpr int BN = Cfg::BN, STAGES = Cfg::STAGES, STAGE_DOUBLES = Cfg::STAGE_DOUBLES;
    constexpr int SLAB_DOUBLES = Cfg::SLAB_A, PITCH_B = Cfg::PITCH_B;
    constexpr int GEMM_CONSUMER_WARPS = Cfg::CONSUMER_WARPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(STAGES) * STAGE_DOUBLES);
    // bars[0..STAGES) = full, bars[STAGES..2*STAGES) = empty

    const int nsub = p.nbm * p.nbn;
    const int task_id = blockIdx.x / nsub;
    const int sub = blockIdx.x - task_id * nsub;
    const int bm = sub % p.nbm;
    const int bn = sub / p.nbm;

    const chol_task_t task = load_task(p, task_id);
    const bool lower = (task.flags & CHOL_TASK_LOWER) != 0;
    const int row0 = bm * BM, col0 = bn * BN;
    if (lower && col0 >= row0 + BM) return;  // block strictly above the diagonal: nothing to do

    const int mv = min(BM, p.m - row0);  // valid rows / cols of this block
    const int nv = min(BN, p.n - col0);
    const int nk = (p.k + BK - 1) / BK;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    pdl_sync();      // (the task records above are plan data, never written by a predecessor kernel)

    // The host checks the 16-byte alignment the bulk copies and the 16-byte C accesses need only for a
    // single task; the pointers of a device task list / tile list are checked here, per task.  A task
    // that is merely 8-byte aligned is computed by this CTA with plain FMAs (slow, correct) instead of
    // raising a misaligned-address fault that would kill the context.  C == A (the in-place multiply)
    // cannot take this path and is rejected on the host when misaligned.
    if (((reinterpret_cast<uintptr_t>(task.C) | reinterpret_cast<uintptr_t>(task.A) |
          reinterpret_cast<uintptr_t>(task.B)) & 15u) != 0) {
        gemm_block_scalar(p, task, lower, row0, col0, mv, nv);
        return;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);                              // producer's expect_tx arrive
            mbar_init(smem_u32(&bars[STAGES + s]), GEMM_CONSUMER_WARPS);   // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == GEMM_CONSUMER_WARPS) {
        // ===================== producer warp: TMA bulk staging =====================
        const double* gA = task.A + row0;
        const double* gB = task.B + col0;
        if (p.beta != 0.0) {
            // warm L2 with this block of C for the epilogue
            const double* gCp = task.C + size_t(col0) * p.ldc + row0;
            for (int c = lane; c < nv; c += 32) bulk_prefetch_l2(gCp + size_t(c) * p.ldc, uint32_t(mv) * 8u);
        }
        const bool isA = lane < BK;
        const int col = isA ? lane : lane - BK;
        const double* gsrc = isA ? gA : gB;
        const int ld = isA ? p.lda : p.ldb;
        const uint32_t bytes = uint32_t(isA ? mv : nv) * 8u;
        for (int it = 0; it < nk; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            const uint32_t full = smem_u32(&bars[s]);
            mbar_wait(smem_u32(&bars[STAGES + s]), ph ^ 1);
            const int k0 = it * BK;
            const int kc = min(BK, p.k - k0);
            if (lane == 0) mbar_expect_tx(full, uint32_t(kc) * uint32_t(mv + nv) * 8u);
            __syncwarp();
            if (col < kc) {
                double* dst = smem + size_t(s) * STAGE_DOUBLES + (isA ? col * PITCH : SLAB_DOUBLES + col * PITCH_B);
                bulk_g2s(smem_u32(dst), gsrc + size_t(k0 + col) * ld, bytes, full);
            }
        }
        return;
    }

    // ========================= consumer warps: DMMA main loop ==========================
    const int wm = warp & 1;   // 2 warps along m (64 rows each)
    const int wn = warp >> 1;  // BN/32 warps along n (32 cols each)
    const int g = lane >> 2;   // mma group id  -> rows 2g, 2g+1 of a 16-row block
    const int t = lane & 3;    // thread in group -> k index / columns 4t..4t+3

    // acc[q][r][mp][np][e]: 16x16 block (q along m, r along n), mma (mp,np), element e
    double acc[4][2][2][2][2];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int mp = 0; mp < 2; ++mp)
#pragma unroll
                for (int np = 0; np < 2; ++np) acc[q][r][mp][np][0] = acc[q][r][mp][np][1] = 0.0;

    const int a_off = t * PITCH + wm * 64 + 2 * g;                  // + q*16 + kk*PITCH
    const int b_off = SLAB_DOUBLES + t * PITCH_B + wn * 32 + 2 * g; // + r*16 + kk*PITCH_B

    for (int it = 0; it < nk; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(smem_u32(&bars[s]), ph);
        const double* st = smem + size_t(s) * STAGE_DOUBLES;
        const int kc = min(BK, p.k - it * BK);
        // one k4 step: 6 LDS.128 feed 32 DMMAs
        auto k4_step = [&](const int kk) {
            double2 a[4], b[2];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                a[q] = *reinterpret_cast<const double2*>(st + a_off + kk * PITCH + q * 16);
#pragma unroll
            for (int r = 0; r < 2; ++r)
                b[r] = *reinterpret_cast<const double2*>(st + b_off + kk * PITCH_B + r * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    dmma884(acc[q][r][0][0][0], acc[q][r][0][0][1], a[q].x, b[r].x);
                    dmma884(acc[q][r][0][1][0], acc[q][r][0][1][1], a[q].x, b[r].y);
                    dmma884(acc[q][r][1][0][0], acc[q][r][1][0][1], a[q].y, b[r].x);
                    dmma884(acc[q][r][1][1][0], acc[q][r][1][1][1], a[q].y, b[r].y);
                }
        };
        // (Specialising full slabs into one branch-free basic block, so that ptxas schedules the fragment loads of
        // a step under the DMMAs of the step before, was measured on B200: 34.27 against 34.38 TFLOP/s for this
        // form at 120 tasks, 34.93 / 34.94 at 820 — the co-resident CTA already fills those bubbles.)
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            if (kk < kc) k4_step(kk);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[STAGES + s]));
    }

    // ================================ epilogue ========================================
    // C <- beta*C + alpha*acc.  The products are summed on their own (from zero) and C enters once at the
    // end: starting the accumulators at C would save the loads below but rounds every one of the K FMAs at
    // the magnitude of C — measured: backward error 3e-15 instead of 4e-16 on the diagonally dominant test
    // matrices.  The eight 16-byte loads of a 16-row group are issued together (the previous one-at-a-time
    // read-modify-write paid 32 dependent round trips per thread: 18 us of a 29 us K = 128 launch); the
    // producer warmed L2 with this block of C while the main loop ran.
    const double alpha = p.alpha, beta = p.beta;
    const bool diag_block = lower && (col0 + BN > row0);   // block touches the diagonal
    double* gC = task.C + size_t(col0) * p.ldc + row0;
    // One group = the 16 rows q, the 16-column halves r_lo .. r_hi-1 and elements e_lo .. e_hi-1 of this warp's
    // sub-tile.  While all 64 accumulators are still live the groups are small (2, 2, then 4 loads in flight),
    // from q == 1 on a group is 8 loads: the full batch from the start costs 120 bytes of spills under the
    // 168-register cap ptxas applies for two 5-warp CTAs per SM (ptxas -v; now 0 spills).
    auto group = [&](const int q, const int r_lo, const int r_hi, const int e_lo = 0, const int e_hi = 2) {
        const int r_loc = wm * 64 + q * 16 + 2 * g;  // first of the two rows this lane owns
        if (r_loc >= mv) return;                     // mv is even: the pair is in or out together
        double2 old[2][2][2];
#pragma unroll
        for (int r = r_lo; r < r_hi; ++r)
#pragma unroll
            for (int e = e_lo; e < e_hi; ++e)
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    const int c_loc = wn * 32 + r * 16 + 4 * t + 2 * e + np;
                    old[r][e][np] = make_double2(0.0, 0.0);
                    if (beta == 0.0 || c_loc >= nv) continue;
                    const double* ptr = gC + size_t(c_loc) * p.ldc + r_loc;
                    if (!diag_block || row0 + r_loc >= col0 + c_loc)
                        old[r][e][np] = *reinterpret_cast<const double2*>(ptr);
                    else if (row0 + r_loc + 1 == col0 + c_loc)
                        old[r][e][np].y = ptr[1];     // the pair straddles the diagonal: lower element only
                }
#pragma unroll
        for (int r = r_lo; r < r_hi; ++r)
#pragma unroll
            for (int e = e_lo; e < e_hi; ++e)
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    const int c_loc = wn * 32 + r * 16 + 4 * t + 2 * e + np;
                    if (c_loc >= nv) continue;
                    double* ptr = gC + size_t(c_loc) * p.ldc + r_loc;
                    double v0 = alpha * acc[q][r][0][np][e];
                    double v1 = alpha * acc[q][r][1][np][e];
                    if (beta != 0.0) {
                        v0 += beta * old[r][e][np].x;
                        v1 += beta * old[r][e][np].y;
                    }
                    if (diag_block && row0 + r_loc < col0 + c_loc) {
                        // pair straddles or lies above the diagonal
                        if (row0 + r_loc + 1 == col0 + c_loc) ptr[1] = v1;
                        continue;
                    }
                    *reinterpret_cast<double2*>(ptr) = make_double2(v0, v1);
                }
    };
    group(0, 0, 1, 0, 1);
    group(0, 0, 1, 1, 2);
    group(0, 1, 2);
    group(1, 0, 2);
    group(2, 0, 2);
    group(3, 0, 2);
}

// ---- generic fallback for shapes the fast path cannot take (odd sizes, unaligned) ------
// Plain CUDA, one thread per C element.  Only ever used for tiny / odd tiles (e.g. the
// client's default B=4, C2:350) where speed is irrelevant; still GPU code, never the CPU.
__global__ void gemm_nt_generic_kernel(const GemmParams p) {
    const int task_id = blockIdx.z;
    const chol_task_t task = load_task(p, task_id);
    const bool lower = (task.flags & CHOL_TASK_LOWER) != 0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= p.m || j >= p.n) return;
    if (lower && i < j) return;
    double s = 0.0;
    for (int l = 0; l < p.k; ++l) s = fma(task.A[size_t(l) * p.lda + i], task.B[size_t(l) * p.ldb + j], s);
    double* c = task.C + size_t(j) * p.ldc + i;
    double v = p.alpha * s;
    if (p.beta != 0.0) v += p.beta * *c;
    *c = v;
}

// In-place right multiplication by the transpose of a lower-triangular matrix, for the shapes
// the fast path cannot take:  A <- A * W^T  (A m x n, W n x n lower).  Column j of the result
// needs columns l <= j of A only, so one thread per row sweeping j downwards is alias-safe.
__global__ void trmm_rlt_inplace_generic_kernel(const GemmParams p) {
    const chol_task_t task = load_task(p, blockIdx.y);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.m) return;
    double* a = task.C + i;
    for (int j = p.n - 1; j >= 0; --j) {
        double s = 0.0;
        for (int l = 0; l <= j && l < p.k; ++l) s = fma(a[size_t(l) * p.ldc], task.B[size_t(l) * p.ldb + j], s);
        a[size_t(j) * p.ldc] = p.alpha * s;
    }
}

}  // namespace chol
