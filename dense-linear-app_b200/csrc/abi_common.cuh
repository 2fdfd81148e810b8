// abi_common.cuh — error reporting and launch accounting shared by the translation units of
// libchol_b200.so.  Nothing here is exported; the C ABI is include/chol_b200.h.
#pragma once
#include <atomic>
#include <string>
#include <cuda_runtime.h>

namespace chol_abi {

extern thread_local std::string g_err;                 // chol_last_error()
extern std::atomic<unsigned long long> g_launches;     // kernels enqueued by this library (chol_launch_count)

inline int fail_cuda(cudaError_t e, const char* where) {
    g_err = std::string(where) + ": " + cudaGetErrorString(e);
    return int(e) > 0 ? int(e) : 1;
}
inline int fail_arg(int idx, const char* fn, const char* what) {
    g_err = std::string(fn) + ": bad argument " + std::to_string(idx) + " (" + what + ")";
    return -idx;
}
// one-time per-device setup (shared-memory attributes); defined in chol_abi.cu
int ensure_init();

}  // namespace chol_abi

#define CHECK_LAUNCH(where)                                                   \
    do {                                                                      \
        cudaError_t e__ = cudaGetLastError();                                 \
        if (e__ != cudaSuccess) return chol_abi::fail_cuda(e__, where);       \
        chol_abi::g_launches.fetch_add(1, std::memory_order_relaxed);         \
    } while (0)
