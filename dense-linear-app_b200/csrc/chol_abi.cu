// chol_abi.cu — the C ABI of libchol_b200.so (see include/chol_b200.h for the contract and
// the reference interfaces each entry point replaces).  Host code here only validates
// arguments and enqueues kernels; there is no CPU arithmetic and no CPU fallback.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "abi_common.cuh"
#include "aux.cuh"
#include "batched.cuh"
#include "batched_np.cuh"
#include "gemm_dmma.cuh"
#include "panel.cuh"
#include "panel2.cuh"

using namespace chol;

namespace chol_abi {
thread_local std::string g_err;
std::atomic<unsigned long long> g_launches{0};
}  // namespace chol_abi
using namespace chol_abi;

namespace {

std::mutex g_mu;
bool g_force_wide = false;  // CHOL_GEMM_WIDE=1: use the 128x128 shape everywhere (A/B experiments)
bool g_inited[64] = {false};
bool g_panel_v1 = false;    // CHOL_PANEL_V1=1: round-1 panel kernels (full 128x128 inverses) for every tile size
bool g_pdl = true;          // CHOL_PDL=0: no programmatic dependent launch for the panel chain
int g_batched_ll = 5;       // CHOL_BATCHED_LL: 5 = left-looking DMMA kernel without a producer warp, 4 CTAs/SM x 128
                            // registers (batched_np.cuh, default); 4 = with a producer warp, 4 CTAs/SM x 96 registers;
                            // 6 = 6 stages x 3 CTAs/SM; 0 = round-1 kernels (A/B experiments)

}  // namespace

namespace chol_abi {
int ensure_init() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDevice (no CUDA device: this library has no CPU fallback)");
    if (dev < 0 || dev >= 64) return fail_arg(1, "chol_init", "device index");
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_inited[dev]) return 0;
    if (const char* w = getenv("CHOL_GEMM_WIDE")) g_force_wide = (w[0] == '1');
    e = cudaFuncSetAttribute(gemm_nt_dmma_kernel<GemmWide>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int(GemmWide::SMEM_BYTES));
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(gemm_nt_dmma_kernel<GemmWide>)");
    e = cudaFuncSetAttribute(gemm_nt_dmma_kernel<GemmPair>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int(120 * 1024));
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(gemm_nt_dmma_kernel<GemmPair>)");
    e = cudaFuncSetAttribute(gemm_nt_dmma_kernel<GemmPair>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(carveout)");
    e = cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(DIAG_SMEM_BYTES));
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(potrf_diag_kernel)");
    e = cudaFuncSetAttribute(trtri_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(DIAG_SMEM_BYTES));
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(trtri_diag_kernel)");
    e = cudaFuncSetAttribute(potrf_batched_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int(DIAG_SMEM_BYTES));
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(potrf_batched_smem_kernel)");
    e = cudaFuncSetAttribute(potrf_batched_global_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int(batched_global_smem(BATCHED_MAX_N)));
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(potrf_batched_global_kernel)");
    e = cudaFuncSetAttribute(potrf_batched_ll_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int(BatchedLL<4>::SMEM));
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(potrf_batched_ll_kernel<4,4>)");
    e = cudaFuncSetAttribute(potrf_batched_ll_kernel<6, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int(BatchedLL<6>::SMEM));
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(potrf_batched_ll_kernel<6,3>)");
    cudaFuncSetAttribute(potrf_batched_ll_kernel<4, 4>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(potrf_batched_ll_kernel<6, 3>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    e = cudaFuncSetAttribute(potrf_batched_np_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int(BatchedNP<4>::SMEM));
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(potrf_batched_np_kernel<4,4>)");
    cudaFuncSetAttribute(potrf_batched_np_kernel<4, 4>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    if (const char* w = getenv("CHOL_BATCHED_LL")) g_batched_ll = atoi(w);
    if (const char* w = getenv("CHOL_PANEL_V1")) g_panel_v1 = (w[0] == '1');
    if (const char* w = getenv("CHOL_PDL")) g_pdl = (w[0] != '0');
    e = cudaFuncSetAttribute(potrf_diag32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(D2_SMEM));
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(potrf_diag32_kernel)");
    e = cudaFuncSetAttribute(trsm_leaf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LF_SMEM));
    if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(trsm_leaf32_kernel)");
    cudaFuncSetAttribute(potrf_diag32_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(trsm_leaf32_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    g_inited[dev] = true;
    return 0;
}
}  // namespace chol_abi

namespace {

constexpr size_t GEMM_THIN_SMEM = 120 * 1024;

// Programmatic dependent launch for the kernels of the panel chain (see pdl_sync in gemm_dmma.cuh): set by
// chol_potrf_tile / the TRSM sweep around their launches.  CHOL_PDL=0 turns it off.
thread_local bool t_chain = false;
struct ChainScope {
    bool prev;
    ChainScope() : prev(t_chain) { t_chain = g_pdl; }
    ~ChainScope() { t_chain = prev; }
};

template <class... KArgs, class... Args>
cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = t_chain ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

GemmParams make_params(int ntasks, int m, int n, int k, int lda, int ldb, int ldc, double alpha, double beta) {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.ntasks = ntasks;
    p.m = m; p.n = n; p.k = k;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.alpha = alpha; p.beta = beta;
    return p;
}

// Enqueue a grouped update.  The task source is p.tasks (device list), p.tile_ptrs (panel mode)
// or p.one (single task, whose pointers are checked for the fast path's 16-byte alignment; the
// caller vouches for the alignment of device lists).
// `inplace_tri`: the tasks have C == A, beta == 0 and B lower triangular (the multiply step of
// the blocked TRSM); the fast path handles the aliasing by giving all n (<= 128) columns of a row
// block to ONE CTA, the generic path has an alias-safe kernel.
int launch_gemm(GemmParams p, cudaStream_t st, bool inplace_tri = false, bool thin = false) {
    const int ntasks = p.ntasks, m = p.m, n = p.n, k = p.k;
    if (ntasks <= 0 || m <= 0 || n <= 0) return 0;
    const bool single = !p.tasks && !p.tile_ptrs;
    bool fast = (m % 2 == 0) && (n % 2 == 0) && (k % 4 == 0) && (k > 0) && (p.lda % 2 == 0) && (p.ldb % 2 == 0) &&
                (p.ldc % 2 == 0) && (p.c_off % 2 == 0) && (p.a_off % 2 == 0);
    if (fast && single) fast = aligned16(p.one.A) && aligned16(p.one.B) && aligned16(p.one.C);
    if (fast && p.tile_ptrs) fast = aligned16(p.one.B);
    if (fast) {
        const bool wide = inplace_tri || g_force_wide;
        const int bn = wide ? GemmWide::BN : GemmPair::BN;
        if (inplace_tri && n > GemmWide::BN) return fail_arg(4, "launch_gemm", "in-place multiply wider than one CTA");
        p.nbm = (m + BM - 1) / BM;
        p.nbn = (n + bn - 1) / bn;
        const long long grid = (long long)ntasks * p.nbm * p.nbn;
        if (grid > 0x7fffffffLL) return fail_arg(2, "chol_gemm_tasks", "too many CTA tiles");
        if (wide)
            launch_k(gemm_nt_dmma_kernel<GemmWide>, dim3((unsigned)grid), dim3(GemmWide::THREADS), GemmWide::SMEM_BYTES, st, p);
        else
            // `thin`: ask for more dynamic shared memory than the kernel uses so that only ONE CTA fits per SM and
            // the other half of every SM stays free for the panel kernels of the next step
            launch_k(gemm_nt_dmma_kernel<GemmPair>, dim3((unsigned)grid), dim3(GemmPair::THREADS),
                     thin ? GEMM_THIN_SMEM : size_t(GemmPair::SMEM_BYTES), st, p);
        CHECK_LAUNCH("gemm_nt_dmma_kernel");
        return 0;
    }
    for (int t0 = 0; t0 < ntasks; t0 += 32768) {
        GemmParams q = p;
        q.ntasks = (ntasks - t0 < 32768) ? ntasks - t0 : 32768;
        if (p.tasks) q.tasks = p.tasks + t0;
        if (p.tile_ptrs) q.tile_ptrs = p.tile_ptrs + t0;
        if (inplace_tri) {
            trmm_rlt_inplace_generic_kernel<<<dim3((m + 127) / 128, q.ntasks), 128, 0, st>>>(q);
            CHECK_LAUNCH("trmm_rlt_inplace_generic_kernel");
        } else {
            gemm_nt_generic_kernel<<<dim3((m + 31) / 32, (n + 7) / 8, q.ntasks), dim3(32, 8), 0, st>>>(q);
            CHECK_LAUNCH("gemm_nt_generic_kernel");
        }
    }
    return 0;
}

int launch_one(const chol_task_t& t, int m, int n, int k, int lda, int ldb, int ldc, double alpha, double beta,
               cudaStream_t st, bool inplace_tri = false) {
    GemmParams p = make_params(1, m, n, k, lda, ldb, ldc, alpha, beta);
    p.one = t;
    return launch_gemm(p, st, inplace_tri);
}

// Tiles whose size is a multiple of 32 take the second-generation panel kernels (panel2.cuh).
inline bool fast32(int b, int lda, int ldl, const void* a, const void* l) {
    return !g_panel_v1 && b % SB == 0 && lda % 2 == 0 && ldl % 2 == 0 && aligned16(a) && aligned16(l);
}

// X = A L_jj^{-T} on the 128-column block at `off` of one matrix (`single`) or of every tile of a
// pointer list: m rows, nbv (multiple of 32) columns.
int launch_leaf(double* const* d_tiles, double* single, int ntiles, long long off, int m, int nbv, int lda,
                const double* Ljj, int ldl, const double* Dinv, cudaStream_t st, const long long* peer_dst = nullptr,
                int npeer = 0) {
    if (ntiles <= 0 || m <= 0 || nbv <= 0) return 0;
    LeafParams p;
    p.tile_ptrs = d_tiles; p.single = single; p.off = off;
    p.m = m; p.nbk = nbv / SB; p.lda = lda;
    p.L = Ljj; p.ldl = ldl; p.Dinv = Dinv;
    p.peer_dst = peer_dst; p.npeer = peer_dst ? npeer : 0;
    p.ctas_per_task = (m + LF_ROWS - 1) / LF_ROWS;
    const long long grid = (long long)ntiles * p.ctas_per_task;
    if (grid > 0x7fffffffLL) return fail_arg(6, "chol_trsm_tiles", "too many CTAs");
    launch_k(trsm_leaf32_kernel, dim3((unsigned)grid), dim3(LF_THREADS), LF_SMEM, st, p);
    CHECK_LAUNCH("trsm_leaf32_kernel");
    return 0;
}

// X * L^T = A for one tile (`single`) or every tile of a panel (`d_tiles`), recursively over the
// 128-column blocks [lo, hi) of L:   solve(lo, mid);  A[:, mid:hi) -= X[:, lo:mid) L[mid:hi, lo:mid)^T;
// solve(mid, hi);   a leaf multiplies by the inverted diagonal block: X_j = A_j inv(L_jj)^T.
// Compared with a left-to-right sweep the dependent chain of GEMM K-lengths drops from
// 128*(1+2+..+7) to 4*128 + 2*256 + 512 and the big updates expose more parallel CTAs.
struct TrsmCtx {
    int b;
    const double* L;
    int ldl;
    const double* Winv;
    double* const* d_tiles;
    double* single;
    int ntiles, lda;
    cudaStream_t st;
    bool v2;        // leaves by block substitution with the 32x32 inverses (trsm_leaf32_kernel)
    const long long* peer_dst;   // fused push of the finished column blocks into peers' slots (v2 only)
    int npeer;
};

int trsm_issue(const TrsmCtx& c, long long c_off, long long a_off, const double* B, int n, int k, int ldb,
               double alpha, double beta, bool inplace) {
    GemmParams p = make_params(c.ntiles, c.b, n, k, c.lda, ldb, c.lda, alpha, beta);
    if (c.single) {
        p.one.C = c.single + c_off; p.one.A = c.single + a_off; p.one.B = B; p.one.flags = 0;
    } else {
        p.tile_ptrs = c.d_tiles; p.c_off = c_off; p.a_off = a_off; p.one.B = B;
    }
    return launch_gemm(p, c.st, inplace);
}

int trsm_rec(const TrsmCtx& c, int lo, int hi) {
    if (hi - lo == 1) {
        const int o = lo * NBD;
        const int nbv = (c.b - o < NBD) ? c.b - o : NBD;
        if (c.v2)
            return launch_leaf(c.d_tiles, c.single, c.ntiles, (long long)o * c.lda, c.b, nbv, c.lda,
                               c.L + size_t(o) * c.ldl + o, c.ldl, c.Winv + size_t(lo) * NBD * NBD, c.st, c.peer_dst,
                               c.npeer);
        return trsm_issue(c, (long long)o * c.lda, (long long)o * c.lda, c.Winv + size_t(lo) * NBD * NBD, nbv, nbv, NBD,
                          1.0, 0.0, true);
    }
    const int mid = lo + (hi - lo + 1) / 2;
    if (int rc = trsm_rec(c, lo, mid)) return rc;
    const int o_lo = lo * NBD, o_mid = mid * NBD;
    const int o_hi = (hi * NBD < c.b) ? hi * NBD : c.b;
    if (int rc = trsm_issue(c, (long long)o_mid * c.lda, (long long)o_lo * c.lda, c.L + size_t(o_lo) * c.ldl + o_mid,
                            o_hi - o_mid, o_mid - o_lo, c.ldl, -1.0, 1.0, false))
        return rc;
    return trsm_rec(c, mid, hi);
}

int trsm_sweep(int b, const double* L, int ldl, const double* Winv, double* const* d_tiles, double* single,
               int ntiles, int lda, cudaStream_t st, const long long* peer_dst = nullptr, int npeer = 0) {
    // (the tiles of a pointer list are 16-byte aligned by contract, see chol_b200.h)
    const bool v2 = fast32(b, lda, ldl, single ? (const void*)single : (const void*)L, L) && aligned16(Winv);
    if (peer_dst && !v2) return fail_arg(1, "chol_trsm_tiles_push", "the fused push needs tiles that are multiples of 32");
    TrsmCtx c{b, L, ldl, Winv, d_tiles, single, ntiles, lda, st, v2, peer_dst, npeer};
    ChainScope chain;
    return trsm_rec(c, 0, (b + NBD - 1) / NBD);
}

}  // namespace

extern "C" {

int chol_init(int device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail_cuda(e, "cudaSetDevice");
    return ensure_init();
}
int chol_finalize(void) { return 0; }
const char* chol_last_error(void) { return g_err.c_str(); }
const char* chol_version(void) { return "chol_b200 0.1 sm_100a"; }
unsigned long long chol_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int chol_gemm_tasks(const chol_task_t* d_tasks, int ntasks, int m, int n, int k, int lda, int ldb, int ldc,
                    double alpha, double beta, void* stream) {
    if (ntasks < 0) return fail_arg(2, "chol_gemm_tasks", "ntasks");
    if (ntasks > 0 && !d_tasks) return fail_arg(1, "chol_gemm_tasks", "d_tasks");
    if (m < 0) return fail_arg(3, "chol_gemm_tasks", "m");
    if (n < 0) return fail_arg(4, "chol_gemm_tasks", "n");
    if (k < 0) return fail_arg(5, "chol_gemm_tasks", "k");
    if (lda < (m > 1 ? m : 1)) return fail_arg(6, "chol_gemm_tasks", "lda");
    if (ldb < (n > 1 ? n : 1)) return fail_arg(7, "chol_gemm_tasks", "ldb");
    if (ldc < (m > 1 ? m : 1)) return fail_arg(8, "chol_gemm_tasks", "ldc");
    if (int rc = ensure_init()) return rc;
    GemmParams p = make_params(ntasks, m, n, k, lda, ldb, ldc, alpha, beta);
    p.tasks = d_tasks;
    return launch_gemm(p, (cudaStream_t)stream);
}

int chol_gemm_tasks_ex(const chol_task_t* d_tasks, int ntasks, int m, int n, int k, int lda, int ldb, int ldc,
                       double alpha, double beta, int ctas_per_sm, void* stream) {
    if (ctas_per_sm != 1 && ctas_per_sm != 2) return fail_arg(11, "chol_gemm_tasks_ex", "ctas_per_sm (1 or 2)");
    if (ntasks < 0) return fail_arg(2, "chol_gemm_tasks_ex", "ntasks");
    if (ntasks > 0 && !d_tasks) return fail_arg(1, "chol_gemm_tasks_ex", "d_tasks");
    if (m < 0 || n < 0 || k < 0) return fail_arg(3, "chol_gemm_tasks_ex", "m / n / k");
    if (lda < (m > 1 ? m : 1)) return fail_arg(6, "chol_gemm_tasks_ex", "lda");
    if (ldb < (n > 1 ? n : 1)) return fail_arg(7, "chol_gemm_tasks_ex", "ldb");
    if (ldc < (m > 1 ? m : 1)) return fail_arg(8, "chol_gemm_tasks_ex", "ldc");
    if (int rc = ensure_init()) return rc;
    GemmParams p = make_params(ntasks, m, n, k, lda, ldb, ldc, alpha, beta);
    p.tasks = d_tasks;
    return launch_gemm(p, (cudaStream_t)stream, false, ctas_per_sm == 1);
}

size_t chol_potrf_tile_workspace(int b) {
    if (b <= 0) return 0;
    return size_t((b + NBD - 1) / NBD) * NBD * NBD * sizeof(double);
}

int chol_potrf_tile(int b, double* A, int lda, double* work, int* d_info, int info_base, void* stream) {
    if (b < 0) return fail_arg(1, "chol_potrf_tile", "b");
    if (b == 0) return 0;
    if (!A) return fail_arg(2, "chol_potrf_tile", "A");
    if (lda < b) return fail_arg(3, "chol_potrf_tile", "lda");
    if (!work) return fail_arg(4, "chol_potrf_tile", "work");
    if (int rc = ensure_init()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = (b + NBD - 1) / NBD;
    const bool v2 = fast32(b, lda, lda, A, A) && aligned16(work);
    ChainScope chain;
    for (int j = 0; j < nblk; ++j) {
        const int o = j * NBD;
        const int nbv = (b - o < NBD) ? b - o : NBD;
        const int rem = b - o - nbv;
        double* Wj = work + size_t(j) * NBD * NBD;
        double* Ajj = A + size_t(o) * lda + o;
        if (v2) {
            launch_k(potrf_diag32_kernel, dim3(1), dim3(D2_THREADS), D2_SMEM, st, nbv, Ajj, lda, Wj, d_info, info_base + o);
            CHECK_LAUNCH("potrf_diag32_kernel");
        } else {
            potrf_diag_kernel<<<1, DIAG_THREADS, DIAG_SMEM_BYTES, st>>>(nbv, Ajj, lda, Wj, d_info, info_base + o);
            CHECK_LAUNCH("potrf_diag_kernel");
        }
        if (rem > 0) {
            chol_task_t t;
            // rows below the diagonal block: X = A * inv(L_jj)^T (in place; each CTA owns its rows)
            int rc;
            if (v2) {
                rc = launch_leaf(nullptr, Ajj + nbv, 1, 0, rem, nbv, lda, Ajj, lda, Wj, st);
            } else {
                t.C = Ajj + nbv; t.A = t.C; t.B = Wj; t.flags = 0;
                rc = launch_one(t, rem, nbv, nbv, lda, NBD, lda, 1.0, 0.0, st, true);
            }
            if (rc) return rc;
            // trailing lower triangle: A22 -= X * X^T
            t.C = A + size_t(o + nbv) * lda + (o + nbv); t.A = Ajj + nbv; t.B = t.A; t.flags = CHOL_TASK_LOWER;
            rc = launch_one(t, rem, rem, nbv, lda, lda, lda, -1.0, 1.0, st);
            if (rc) return rc;
        }
    }
    return 0;
}

size_t chol_trsm_tile_workspace(int b) { return chol_potrf_tile_workspace(b); }

int chol_trsm_tile(int b, const double* L, int ldl, double* A, int lda, double* work, void* stream) {
    if (b < 0) return fail_arg(1, "chol_trsm_tile", "b");
    if (b == 0) return 0;
    if (!L) return fail_arg(2, "chol_trsm_tile", "L");
    if (ldl < b) return fail_arg(3, "chol_trsm_tile", "ldl");
    if (!A) return fail_arg(4, "chol_trsm_tile", "A");
    if (lda < b) return fail_arg(5, "chol_trsm_tile", "lda");
    if (!work) return fail_arg(6, "chol_trsm_tile", "work");
    if (int rc = ensure_init()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = (b + NBD - 1) / NBD;
    trtri_diag_kernel<<<nblk, DIAG_THREADS, DIAG_SMEM_BYTES, st>>>(b, L, ldl, work);
    CHECK_LAUNCH("trtri_diag_kernel");
    return trsm_sweep(b, L, ldl, work, nullptr, A, 1, lda, st);
}

int chol_trsm_tiles(int b, const double* L, int ldl, const double* potrf_work, double* const* d_tiles, int ntiles,
                    int lda, void* d_task_scratch, void* stream) {
    if (b < 0) return fail_arg(1, "chol_trsm_tiles", "b");
    if (ntiles < 0) return fail_arg(6, "chol_trsm_tiles", "ntiles");
    if (b == 0 || ntiles == 0) return 0;
    if (!L) return fail_arg(2, "chol_trsm_tiles", "L");
    if (ldl < b) return fail_arg(3, "chol_trsm_tiles", "ldl");
    if (!potrf_work) return fail_arg(4, "chol_trsm_tiles", "potrf_work");
    if (!d_tiles) return fail_arg(5, "chol_trsm_tiles", "d_tiles");
    if (lda < b) return fail_arg(7, "chol_trsm_tiles", "lda");
    if (int rc = ensure_init()) return rc;
    (void)d_task_scratch;
    return trsm_sweep(b, L, ldl, potrf_work, d_tiles, nullptr, ntiles, lda, (cudaStream_t)stream);
}

int chol_trsm_tiles_push(int b, const double* L, int ldl, const double* potrf_work, double* const* d_tiles, int ntiles,
                         int lda, const long long* d_peer_dst, int npeer, void* stream) {
    if (b < 0) return fail_arg(1, "chol_trsm_tiles_push", "b");
    if (ntiles < 0) return fail_arg(6, "chol_trsm_tiles_push", "ntiles");
    if (npeer < 0 || npeer > LF_MAX_PEERS) return fail_arg(9, "chol_trsm_tiles_push", "npeer (at most 7)");
    if (b == 0 || ntiles == 0) return 0;
    if (!L) return fail_arg(2, "chol_trsm_tiles_push", "L");
    if (ldl < b) return fail_arg(3, "chol_trsm_tiles_push", "ldl");
    if (!potrf_work) return fail_arg(4, "chol_trsm_tiles_push", "potrf_work");
    if (!d_tiles) return fail_arg(5, "chol_trsm_tiles_push", "d_tiles");
    if (lda < b) return fail_arg(7, "chol_trsm_tiles_push", "lda");
    if (npeer > 0 && !d_peer_dst) return fail_arg(8, "chol_trsm_tiles_push", "d_peer_dst");
    if (int rc = ensure_init()) return rc;
    return trsm_sweep(b, L, ldl, potrf_work, d_tiles, nullptr, ntiles, lda, (cudaStream_t)stream,
                      npeer > 0 ? d_peer_dst : nullptr, npeer);
}

int chol_syrk_tile(int b, const double* A, int lda, double* C, int ldc, void* stream) {
    if (b < 0) return fail_arg(1, "chol_syrk_tile", "b");
    if (b == 0) return 0;
    if (!A) return fail_arg(2, "chol_syrk_tile", "A");
    if (lda < b) return fail_arg(3, "chol_syrk_tile", "lda");
    if (!C) return fail_arg(4, "chol_syrk_tile", "C");
    if (ldc < b) return fail_arg(5, "chol_syrk_tile", "ldc");
    if (int rc = ensure_init()) return rc;
    chol_task_t t;
    t.C = C; t.A = A; t.B = A; t.flags = CHOL_TASK_LOWER;
    return launch_one(t, b, b, b, lda, lda, ldc, -1.0, 1.0, (cudaStream_t)stream);
}

int chol_gemm_tile(int b, const double* Ai, int ldai, const double* Aj, int ldaj, double* C, int ldc, void* stream) {
    if (b < 0) return fail_arg(1, "chol_gemm_tile", "b");
    if (b == 0) return 0;
    if (!Ai) return fail_arg(2, "chol_gemm_tile", "Ai");
    if (ldai < b) return fail_arg(3, "chol_gemm_tile", "ldai");
    if (!Aj) return fail_arg(4, "chol_gemm_tile", "Aj");
    if (ldaj < b) return fail_arg(5, "chol_gemm_tile", "ldaj");
    if (!C) return fail_arg(6, "chol_gemm_tile", "C");
    if (ldc < b) return fail_arg(7, "chol_gemm_tile", "ldc");
    if (int rc = ensure_init()) return rc;
    chol_task_t t;
    t.C = C; t.A = Ai; t.B = Aj; t.flags = 0;
    return launch_one(t, b, b, b, ldai, ldaj, ldc, -1.0, 1.0, (cudaStream_t)stream);
}

int chol_potrf_batched(int n, int batch, double* A, int lda, long long stride, int* d_info, void* stream) {
    if (n < 0) return fail_arg(1, "chol_potrf_batched", "n");
    if (batch < 0) return fail_arg(2, "chol_potrf_batched", "batch");
    if (n == 0 || batch == 0) return 0;
    if (!A) return fail_arg(3, "chol_potrf_batched", "A");
    if (lda < n) return fail_arg(4, "chol_potrf_batched", "lda");
    if (stride < (long long)lda * (n - 1) + n) return fail_arg(5, "chol_potrf_batched", "stride");
    if (!d_info) return fail_arg(6, "chol_potrf_batched", "d_info");
    if (int rc = ensure_init()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (g_batched_ll != 0 && n % BLW == 0 && n <= BATCHED_LL_MAX_N && lda % 2 == 0 && stride % 2 == 0 &&
        aligned16(A)) {
        if (g_batched_ll == 5)
            potrf_batched_np_kernel<4, 4><<<batch, NP_THREADS, BatchedNP<4>::SMEM, st>>>(n, A, lda, stride, d_info);
        else if (g_batched_ll == 6)
            potrf_batched_ll_kernel<6, 3><<<batch, BL_THREADS, BatchedLL<6>::SMEM, st>>>(n, A, lda, stride, d_info);
        else
            potrf_batched_ll_kernel<4, 4><<<batch, BL_THREADS, BatchedLL<4>::SMEM, st>>>(n, A, lda, stride, d_info);
        CHECK_LAUNCH("potrf_batched_ll_kernel");
        return 0;
    }
    if (n <= NBD) {
        potrf_batched_smem_kernel<<<batch, DIAG_THREADS, DIAG_SMEM_BYTES, st>>>(n, A, lda, stride, d_info);
        CHECK_LAUNCH("potrf_batched_smem_kernel");
    } else {
        if (n > BATCHED_MAX_N) return fail_arg(1, "chol_potrf_batched", "n > 880: use chol_potrf_tile");
        potrf_batched_global_kernel<<<batch, BATCHED_GLOBAL_THREADS, batched_global_smem(n), st>>>(n, A, lda, stride,
                                                                                                 d_info);
        CHECK_LAUNCH("potrf_batched_global_kernel");
    }
    return 0;
}

int chol_plgsy_tile(double bump, int mb, int nb, double* A, int lda, long long bigM, long long row0, long long col0,
                    long long N, unsigned long long seed, void* stream) {
    if (mb < 0) return fail_arg(2, "chol_plgsy_tile", "mb");
    if (nb < 0) return fail_arg(3, "chol_plgsy_tile", "nb");
    if (mb == 0 || nb == 0) return 0;
    if (!A) return fail_arg(4, "chol_plgsy_tile", "A");
    if (lda < mb) return fail_arg(5, "chol_plgsy_tile", "lda");
    if (int rc = ensure_init()) return rc;
    const int threads = 128;
    const int rows_per_block = threads * PLGSY_ROWS_PER_THREAD;
    dim3 grd((mb + rows_per_block - 1) / rows_per_block, nb);
    plgsy_tile_kernel<<<grd, threads, 0, (cudaStream_t)stream>>>(bump, mb, nb, A, lda, (unsigned long long)bigM, row0,
                                                                col0, N, seed);
    CHECK_LAUNCH("plgsy_tile_kernel");
    return 0;
}

int chol_tile_sumsq(int m, int n, const double* A, int lda, int mode, double* d_out, void* stream) {
    if (m <= 0 || n <= 0) return 0;
    if (!A) return fail_arg(3, "chol_tile_sumsq", "A");
    if (!d_out) return fail_arg(6, "chol_tile_sumsq", "d_out");
    if (int rc = ensure_init()) return rc;
    const int threads = 256;
    tile_col_sumsq_kernel<<<(n * 32 + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(m, n, A, lda, mode,
                                                                                                  d_out);
    CHECK_LAUNCH("tile_col_sumsq_kernel");
    return 0;
}

int chol_tile_abs_sums(int m, int n, const double* A, int lda, int mode, double* d_rows, double* d_cols,
                       void* stream) {
    if (m <= 0 || n <= 0) return 0;
    if (!A) return fail_arg(3, "chol_tile_abs_sums", "A");
    if (int rc = ensure_init()) return rc;
    const int threads = 256;
    if (d_cols) {
        tile_abs_colsum_kernel<<<(n * 32 + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(m, n, A, lda,
                                                                                                       mode, d_cols);
        CHECK_LAUNCH("tile_abs_colsum_kernel");
    }
    if (d_rows) {
        tile_abs_rowsum_kernel<<<(m + 127) / 128, 128, 0, (cudaStream_t)stream>>>(m, n, A, lda, mode, d_rows);
        CHECK_LAUNCH("tile_abs_rowsum_kernel");
    }
    return 0;
}

int chol_tile_tril(int n, const double* A, int lda, double* B, int ldb, void* stream) {
    if (n <= 0) return 0;
    if (!A) return fail_arg(2, "chol_tile_tril", "A");
    if (!B) return fail_arg(4, "chol_tile_tril", "B");
    if (int rc = ensure_init()) return rc;
    dim3 grd((n + 127) / 128, n);
    tile_tril_kernel<<<grd, 128, 0, (cudaStream_t)stream>>>(n, A, lda, B, ldb);
    CHECK_LAUNCH("tile_tril_kernel");
    return 0;
}

int chol_tile_transpose(int n, double* A, int lda, long long stride, int ntiles, void* stream) {
    if (n < 0) return fail_arg(1, "chol_tile_transpose", "n");
    if (ntiles < 0) return fail_arg(5, "chol_tile_transpose", "ntiles");
    if (n == 0 || ntiles == 0) return 0;
    if (!A) return fail_arg(2, "chol_tile_transpose", "A");
    if (lda < n) return fail_arg(3, "chol_tile_transpose", "lda");
    if (ntiles > 1 && stride < (long long)lda * n) return fail_arg(4, "chol_tile_transpose", "stride");
    if (int rc = ensure_init()) return rc;
    const int nb = (n + 31) / 32;
    for (int t0 = 0; t0 < ntiles; t0 += 65535) {
        const int cnt = (ntiles - t0 < 65535) ? ntiles - t0 : 65535;
        tile_transpose_kernel<<<dim3(nb * (nb + 1) / 2, cnt), 256, 0, (cudaStream_t)stream>>>(
            n, A + size_t(t0) * size_t(stride), lda, stride);
        CHECK_LAUNCH("tile_transpose_kernel");
    }
    return 0;
}

int chol_fp64_peak(int kind, int iters, double* flops_out, void* stream) {
    if (!flops_out) return fail_arg(3, "chol_fp64_peak", "flops_out");
    if (int rc = ensure_init()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_out, 8);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double flops = 0;
    for (int rep = 0; rep < 2; ++rep) {  // rep 0 = warm-up
        cudaEventRecord(e0, st);
        if (kind == 0) {
            peak_dfma_kernel<<<sms, 1024, 0, st>>>(iters, d_out);
            flops = double(sms) * 1024 * 16 * 2.0 * iters;
        } else {
            const int warps = kind == 1 ? 8 : 4;
            peak_dmma_kernel<<<sms, warps * 32, 0, st>>>(iters, d_out);
            flops = double(sms) * warps * 32 * 512.0 * iters;
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) break;
        cudaEventRecord(e1, st);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
    }
    float ms = 0;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail_cuda(e, "chol_fp64_peak");
    *flops_out = flops / (double(ms) * 1e-3);
    return 0;
}

#ifdef CHOL_DIAG_CLOCKS
// debug builds only: phase timestamps (SM clock) of the most recent potrf_diag32_kernel launch
int chol_debug_diag_clocks(long long* out16) {
    cudaError_t e = cudaMemcpyFromSymbol(out16, g_diag_clk, 16 * sizeof(long long));
    return e == cudaSuccess ? 0 : fail_cuda(e, "chol_debug_diag_clocks");
}
int chol_debug_batched_clocks(long long* out32) {
    cudaError_t e = cudaMemcpyFromSymbol(out32, g_bat_clk, 32 * sizeof(long long));
    return e == cudaSuccess ? 0 : fail_cuda(e, "chol_debug_batched_clocks");
}
#endif

}  // extern "C"
