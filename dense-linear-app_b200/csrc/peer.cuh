// peer.cuh — the panel transport between the GPUs of one box without a resident kernel.
//
// The factored panel of step k (and L_kk with its inverted diagonal blocks) has to reach the
// ranks whose trailing update reads it (2D block-cyclic; the reference carries the grid
// as p,q of CHAMELEON_Desc_Create, V6:44-45, and leaves the transfers to StarPU/MPI).  Round 1 used
// ncclBroadcast: the receivers' NCCL kernels sat on SMs spinning until the owner had factored the
// panel (update kernel 33.3 -> 28.4 TFLOP/s per GPU on 8 GPUs).  Here instead:
//   * receive slots live in cudaMalloc'ed buffers exported with CUDA IPC, every rank maps every
//     peer's buffer (NVLink/NVSwitch P2P);
//   * the owner PUSHES with the copy engines (cudaMemcpyAsync / cudaMemcpy2DAsync on a per-peer send
//     stream), then raises a flag word in the peer's memory with a one-warp kernel (a system-scope
//     release store; the copy before it is complete in stream order);
//   * the reader waits with a stream memory operation (cuStreamWaitValue32, GEQ) — no kernel is
//     resident while it waits, nothing spins on an SM;
//   * slot reuse is guarded by credits flowing the other way through the same flag mechanism.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace chol {

constexpr int PEER_MAX_POST = 16;

struct FlagPost {
    uint32_t* dst[PEER_MAX_POST];
    int n;
    uint32_t value;
};

// thread t: *dst[t] = value, visible system-wide after everything this stream did before.
__global__ void flag_post_kernel(const FlagPost p) {
    const int t = threadIdx.x;
    if (t < p.n) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.dst[t]), "r"(p.value) : "memory");
    }
}

// Fallback wait (CHOL_FLAG_WAIT=kernel): one thread polls until (int)(*flag - value) >= 0.  One rank
// per GPU only — the flag is raised by ANOTHER GPU, never by a kernel queued on this one.
__global__ void flag_spin_kernel(const uint32_t* flag, uint32_t value) {
    uint32_t v;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (int32_t(v - value) >= 0) break;
        __nanosleep(200);
    } while (true);
}

}  // namespace chol
