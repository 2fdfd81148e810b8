// peer_abi.cu — C ABI of the copy-engine panel transport (peer.cuh): IPC-shared receive slots,
// pushes on per-peer send streams, flag words raised by a one-warp kernel and awaited with stream
// memory operations.  See include/chol_b200.h ("panel transport").
#include <cstdlib>
#include <cstring>
#include <mutex>

#include <cuda.h>

#include "../../include/chol_b200.h"
#include "abi_common.cuh"
#include "peer.cuh"

using namespace chol;
using namespace chol_abi;

namespace {

typedef CUresult (*wait32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*write32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

std::mutex g_pmu;
wait32_fn g_wait32 = nullptr;
write32_fn g_write32 = nullptr;
bool g_memops_resolved = false;
int g_wait_mode = 0;   // 0 = stream memory operation, 1 = polling kernel (CHOL_FLAG_WAIT=kernel)
int g_post_mode = 0;   // 0 = kernel store,            1 = stream memory operation (CHOL_FLAG_POST=memop)
cudaEvent_t g_ready[64] = {nullptr};

int resolve_memops() {
    std::lock_guard<std::mutex> lk(g_pmu);
    if (g_memops_resolved) return 0;
    if (const char* w = getenv("CHOL_FLAG_WAIT")) g_wait_mode = (strcmp(w, "kernel") == 0) ? 1 : 0;
    if (const char* w = getenv("CHOL_FLAG_POST")) g_post_mode = (strcmp(w, "memop") == 0) ? 1 : 0;
    // the driver entry points are looked up at run time so the library still loads (symbol checks,
    // argument validation) on a machine without libcuda
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    cudaError_t e = cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) g_wait32 = reinterpret_cast<wait32_fn>(fn);
    fn = nullptr;
    e = cudaGetDriverEntryPoint("cuStreamWriteValue32", &fn, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) g_write32 = reinterpret_cast<write32_fn>(fn);
    (void)cudaGetLastError();
    if (!g_wait32) g_wait_mode = 1;
    if (!g_write32) g_post_mode = 0;
    g_memops_resolved = true;
    return 0;
}

int fail_cu(CUresult r, const char* where) {
    g_err = std::string(where) + ": CUDA driver error " + std::to_string(int(r));
    return 1000 + int(r);
}

int post_flags(const FlagPost& fp, cudaStream_t st) {
    if (fp.n <= 0) return 0;
    if (g_post_mode == 1) {
        for (int i = 0; i < fp.n; ++i) {
            CUresult r = g_write32((CUstream)st, (CUdeviceptr)(uintptr_t)fp.dst[i], fp.value, 0);
            if (r != CUDA_SUCCESS) return fail_cu(r, "cuStreamWriteValue32");
        }
        return 0;
    }
    flag_post_kernel<<<1, 32, 0, st>>>(fp);
    CHECK_LAUNCH("flag_post_kernel");
    return 0;
}

int wait_flag(const uint32_t* flag, uint32_t value, cudaStream_t st) {
    if (g_wait_mode == 0) {
        CUresult r = g_wait32((CUstream)st, (CUdeviceptr)(uintptr_t)flag, value, CU_STREAM_WAIT_VALUE_GEQ);
        if (r != CUDA_SUCCESS) return fail_cu(r, "cuStreamWaitValue32");
        return 0;
    }
    flag_spin_kernel<<<1, 1, 0, st>>>(flag, value);
    CHECK_LAUNCH("flag_spin_kernel");
    return 0;
}

}  // namespace

// ---- SM partition for the panel chain (green contexts) -----------------------------------------
namespace {

struct Partition {
    CUgreenCtx panel_ctx = nullptr, rest_ctx = nullptr;
    CUstream panel_stream = nullptr, rest_stream = nullptr;
};

template <class F>
bool entry(const char* name, F& fn) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
    (void)cudaGetLastError();
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
    fn = reinterpret_cast<F>(p);
    return true;
}

}  // namespace

extern "C" {

int chol_peer_alloc(size_t bytes, void** out) {
    if (!out) return fail_arg(2, "chol_peer_alloc", "out");
    if (int rc = ensure_init()) return rc;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 8);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc (peer buffer)");
    e = cudaMemset(p, 0, bytes ? bytes : 8);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMemset (peer buffer)");
    *out = p;
    return 0;
}

int chol_peer_free(void* p) {
    if (!p) return 0;
    cudaError_t e = cudaFree(p);
    return e == cudaSuccess ? 0 : fail_cuda(e, "cudaFree (peer buffer)");
}

int chol_peer_export(void* p, void* handle64) {
    if (!p) return fail_arg(1, "chol_peer_export", "p");
    if (!handle64) return fail_arg(2, "chol_peer_export", "handle64");
    static_assert(sizeof(cudaIpcMemHandle_t) == CHOL_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) return fail_cuda(e, "cudaIpcGetMemHandle");
    memcpy(handle64, &h, sizeof(h));
    return 0;
}

int chol_peer_open(const void* handle64, void** out) {
    if (!handle64) return fail_arg(1, "chol_peer_open", "handle64");
    if (!out) return fail_arg(2, "chol_peer_open", "out");
    if (int rc = ensure_init()) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail_cuda(e, "cudaIpcOpenMemHandle");
    *out = p;
    return 0;
}

int chol_peer_close(void* p) {
    if (!p) return 0;
    cudaError_t e = cudaIpcCloseMemHandle(p);
    return e == cudaSuccess ? 0 : fail_cuda(e, "cudaIpcCloseMemHandle");
}

int chol_peer_send(const chol_xfer_t* xf, int n, void* ready_stream) {
    if (n < 0) return fail_arg(2, "chol_peer_send", "n");
    if (n == 0) return 0;
    if (!xf) return fail_arg(1, "chol_peer_send", "xf");
    if (int rc = ensure_init()) return rc;
    resolve_memops();
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t e;
    {
        std::lock_guard<std::mutex> lk(g_pmu);
        if (!g_ready[dev]) {
            e = cudaEventCreateWithFlags(&g_ready[dev], cudaEventDisableTiming);
            if (e != cudaSuccess) return fail_cuda(e, "cudaEventCreate");
        }
    }
    // everything the producer stream has enqueued so far (the TRSM of the panel) comes first
    e = cudaEventRecord(g_ready[dev], (cudaStream_t)ready_stream);
    if (e != cudaSuccess) return fail_cuda(e, "cudaEventRecord");
    for (int i = 0; i < n; ++i) {
        const chol_xfer_t& x = xf[i];
        cudaStream_t st = (cudaStream_t)x.stream;
        if (x.count < 0 || x.tile_bytes < 0) return fail_arg(1, "chol_peer_send", "count / tile_bytes");
        e = cudaStreamWaitEvent(st, g_ready[dev], 0);
        if (e != cudaSuccess) return fail_cuda(e, "cudaStreamWaitEvent");
        if (x.credit) {
            if (int rc = wait_flag(x.credit, x.credit_value, st)) return rc;
        }
        if (x.count > 0 && x.tile_bytes > 0) {
            if (!x.dst || !x.src) return fail_arg(1, "chol_peer_send", "dst / src");
            const size_t tb = size_t(x.tile_bytes);
            if (x.count == 1 || (x.dst_stride == 1 && x.src_stride == 1)) {
                e = cudaMemcpyAsync(x.dst, x.src, tb * size_t(x.count), cudaMemcpyDefault, st);
            } else {
                e = cudaMemcpy2DAsync(x.dst, tb * size_t(x.dst_stride), x.src, tb * size_t(x.src_stride), tb,
                                      size_t(x.count), cudaMemcpyDefault, st);
                if (e != cudaSuccess) {   // pitch beyond what the 2-D path takes: tile by tile
                    (void)cudaGetLastError();
                    e = cudaSuccess;
                    for (int t = 0; t < x.count && e == cudaSuccess; ++t)
                        e = cudaMemcpyAsync(static_cast<char*>(x.dst) + tb * size_t(x.dst_stride) * t,
                                            static_cast<const char*>(x.src) + tb * size_t(x.src_stride) * t, tb,
                                            cudaMemcpyDefault, st);
                }
            }
            if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpyAsync (peer push)");
        }
        if (x.flag) {
            FlagPost fp;
            fp.n = 1;
            fp.dst[0] = x.flag;
            fp.value = x.flag_value;
            if (int rc = post_flags(fp, st)) return rc;
        }
    }
    return 0;
}

int chol_flag_wait(const uint32_t* flag, uint32_t value, void* stream) {
    if (!flag) return fail_arg(1, "chol_flag_wait", "flag");
    if (int rc = ensure_init()) return rc;
    resolve_memops();
    return wait_flag(flag, value, (cudaStream_t)stream);
}

int chol_flag_post(uint32_t* const* flags, int n, uint32_t value, void* stream) {
    if (n < 0 || n > PEER_MAX_POST) return fail_arg(2, "chol_flag_post", "n (at most 16 flags per call)");
    if (n == 0) return 0;
    if (!flags) return fail_arg(1, "chol_flag_post", "flags");
    if (int rc = ensure_init()) return rc;
    resolve_memops();
    FlagPost fp;
    fp.n = n;
    fp.value = value;
    for (int i = 0; i < n; ++i) {
        if (!flags[i]) return fail_arg(1, "chol_flag_post", "null flag pointer");
        fp.dst[i] = flags[i];
    }
    return post_flags(fp, (cudaStream_t)stream);
}

int chol_partition_create(int device, int min_panel_sms, chol_partition_t* out) {
    if (!out) return fail_arg(3, "chol_partition_create", "out");
    memset(out, 0, sizeof(*out));
    if (min_panel_sms <= 0) return fail_arg(2, "chol_partition_create", "min_panel_sms");
    if (int rc = ensure_init()) return rc;
    CUresult (*p_devget)(CUdevice*, int) = nullptr;
    CUresult (*p_getres)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
    CUresult (*p_split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int,
                        unsigned int) = nullptr;
    CUresult (*p_desc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
    CUresult (*p_gcreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*p_gdestroy)(CUgreenCtx) = nullptr;
    CUresult (*p_gstream)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
    if (!entry("cuDeviceGet", p_devget) || !entry("cuDeviceGetDevResource", p_getres) ||
        !entry("cuDevSmResourceSplitByCount", p_split) || !entry("cuDevResourceGenerateDesc", p_desc) ||
        !entry("cuGreenCtxCreate", p_gcreate) || !entry("cuGreenCtxDestroy", p_gdestroy) ||
        !entry("cuGreenCtxStreamCreate", p_gstream)) {
        g_err = "chol_partition_create: this driver has no green-context entry points";
        return 1;
    }
    CUdevice dev;
    CUresult r = p_devget(&dev, device);
    if (r != CUDA_SUCCESS) return fail_cu(r, "cuDeviceGet");
    CUdevResource all, grp, rest;
    r = p_getres(dev, &all, CU_DEV_RESOURCE_TYPE_SM);
    if (r != CUDA_SUCCESS) return fail_cu(r, "cuDeviceGetDevResource");
    unsigned int ngroups = 1;
    r = p_split(&grp, &ngroups, &all, &rest, 0, (unsigned int)min_panel_sms);
    if (r != CUDA_SUCCESS) return fail_cu(r, "cuDevSmResourceSplitByCount");
    if (ngroups != 1 || grp.sm.smCount == 0 || rest.sm.smCount == 0) {
        g_err = "chol_partition_create: the device's SMs cannot be split that way";
        return 1;
    }
    Partition* P = new Partition;
    CUdevResourceDesc d1 = nullptr, d2 = nullptr;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    do {
        if ((r = p_desc(&d1, &grp, 1)) != CUDA_SUCCESS) break;
        if ((r = p_desc(&d2, &rest, 1)) != CUDA_SUCCESS) break;
        if ((r = p_gcreate(&P->panel_ctx, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM)) != CUDA_SUCCESS) break;
        if ((r = p_gcreate(&P->rest_ctx, d2, dev, CU_GREEN_CTX_DEFAULT_STREAM)) != CUDA_SUCCESS) break;
        if ((r = p_gstream(&P->panel_stream, P->panel_ctx, CU_STREAM_NON_BLOCKING, prio_hi)) != CUDA_SUCCESS) break;
        if ((r = p_gstream(&P->rest_stream, P->rest_ctx, CU_STREAM_NON_BLOCKING, prio_lo)) != CUDA_SUCCESS) break;
    } while (0);
    if (r != CUDA_SUCCESS) {
        if (P->panel_stream) cudaStreamDestroy((cudaStream_t)P->panel_stream);
        if (P->rest_stream) cudaStreamDestroy((cudaStream_t)P->rest_stream);
        if (P->panel_ctx) p_gdestroy(P->panel_ctx);
        if (P->rest_ctx) p_gdestroy(P->rest_ctx);
        delete P;
        return fail_cu(r, "chol_partition_create (green context setup)");
    }
    out->handle = P;
    out->panel_stream = P->panel_stream;
    out->rest_stream = P->rest_stream;
    out->panel_sms = (int)grp.sm.smCount;
    out->rest_sms = (int)rest.sm.smCount;
    return 0;
}

int chol_partition_destroy(void* handle) {
    if (!handle) return 0;
    Partition* P = static_cast<Partition*>(handle);
    CUresult (*p_gdestroy)(CUgreenCtx) = nullptr;
    if (P->panel_stream) cudaStreamDestroy((cudaStream_t)P->panel_stream);
    if (P->rest_stream) cudaStreamDestroy((cudaStream_t)P->rest_stream);
    if (entry("cuGreenCtxDestroy", p_gdestroy)) {
        if (P->panel_ctx) p_gdestroy(P->panel_ctx);
        if (P->rest_ctx) p_gdestroy(P->rest_ctx);
    }
    delete P;
    return 0;
}

}  // extern "C"
