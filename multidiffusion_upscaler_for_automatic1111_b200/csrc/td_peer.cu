// Multi-GPU plumbing for the tile-sharded sampler step (one process per GPU):
//   * exchange buffers that every rank maps from its peers through CUDA IPC (NVLink 5 / NVSwitch P2P),
//   * a step-counter signal written straight into the peers' flag arrays,
// so that the blend kernel can wait for its peers and read their tile outputs over NVLink in the
// SAME kernel that blends them (td_blend_multidiffusion_peer in td_diffusion.cu): no NCCL call, no
// gathered copy.  The reference has no multi-GPU path at all.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "td_b200.h"
#include "td_internal.h"

namespace {

struct SignalTable {
    uint32_t* ptrs[TD_MAX_PEERS];
};

// One warp: bump this rank's device-side step counter, then publish the new value into slot `rank`
// of every rank's flag array (release, system scope).  Keeping the counter on the device makes the
// launch replayable inside a CUDA graph (no host-supplied step number in the parameters).
__global__ void peer_signal_kernel(const __grid_constant__ SignalTable tbl, int world, int rank, uint32_t* step_counter) {
    const int t = threadIdx.x;
    uint32_t v = 0;
    if (t == 0) {
        v = *step_counter + 1u;
        *step_counter = v;
    }
    v = __shfl_sync(0xffffffffu, v, 0);
    if (t >= world) return;
    __threadfence_system();  // everything this GPU wrote before the signal is visible system-wide first
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(tbl.ptrs[t] + rank), "r"(v) : "memory");
}

// One warp: lane i spins until flags[i] >= *value (acquire, system scope).  Stream-ordered work after this kernel sees
// everything the signalling ranks wrote before their td_peer_signal.  Bounded spin: a missing peer traps (launch error).
__global__ void peer_wait_kernel(const uint32_t* flags, int count, const uint32_t* value) {
    const int t = threadIdx.x;
    if (t >= count) return;
    const uint32_t want = *value;
    uint32_t v;
    unsigned long long spins = 0;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + t) : "memory");
        if (++spins > (1ull << 31)) __trap();
    } while ((int32_t)(v - want) < 0);
}

int cuda_fail(const char* what, cudaError_t e) {
    td_set_error("%s: %s", what, cudaGetErrorString(e));
    return TD_ERR_CUDA;
}

}  // namespace

extern "C" int td_dev_alloc(int64_t bytes, void** out) {
    if (bytes <= 0 || out == nullptr) { td_set_error("td_dev_alloc: bad arguments"); return TD_ERR_INVALID_ARG; }
    cudaError_t e = cudaMalloc(out, (size_t)bytes);
    if (e != cudaSuccess) return cuda_fail("td_dev_alloc", e);
    e = cudaMemset(*out, 0, (size_t)bytes);
    if (e != cudaSuccess) return cuda_fail("td_dev_alloc (memset)", e);
    return TD_OK;
}

extern "C" int td_dev_free(void* ptr) {
    if (ptr == nullptr) return TD_OK;
    cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? TD_OK : cuda_fail("td_dev_free", e);
}

extern "C" int td_ipc_get_handle(const void* dev_ptr, void* handle_out) {
    if (dev_ptr == nullptr || handle_out == nullptr) { td_set_error("td_ipc_get_handle: null"); return TD_ERR_INVALID_ARG; }
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr));
    if (e != cudaSuccess) return cuda_fail("td_ipc_get_handle", e);
    static_assert(sizeof(h) == TD_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    std::memcpy(handle_out, &h, sizeof(h));
    return TD_OK;
}

extern "C" int td_ipc_open(const void* handle, void** dev_ptr_out) {
    if (handle == nullptr || dev_ptr_out == nullptr) { td_set_error("td_ipc_open: null"); return TD_ERR_INVALID_ARG; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? TD_OK : cuda_fail("td_ipc_open", e);
}

extern "C" int td_ipc_close(void* dev_ptr) {
    if (dev_ptr == nullptr) return TD_OK;
    cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
    return e == cudaSuccess ? TD_OK : cuda_fail("td_ipc_close", e);
}

extern "C" int td_peer_signal(void* const* flag_ptrs, int world, int rank, uint32_t* step_counter, void* stream) {
    if (flag_ptrs == nullptr || step_counter == nullptr || world <= 0 || world > TD_MAX_PEERS || rank < 0 || rank >= world) {
        td_set_error("td_peer_signal: bad arguments (world=%d rank=%d)", world, rank);
        return TD_ERR_INVALID_ARG;
    }
    SignalTable tbl;
    for (int i = 0; i < world; ++i) {
        if (flag_ptrs[i] == nullptr) { td_set_error("td_peer_signal: flag_ptrs[%d] is null", i); return TD_ERR_INVALID_ARG; }
        tbl.ptrs[i] = (uint32_t*)flag_ptrs[i];
    }
    peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(tbl, world, rank, step_counter);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? TD_OK : cuda_fail("td_peer_signal", e);
}

extern "C" int td_peer_wait(const uint32_t* flags, int count, const uint32_t* value, void* stream) {
    if (count < 0 || count > TD_MAX_PEERS || (count > 0 && (flags == nullptr || value == nullptr))) {
        td_set_error("td_peer_wait: bad arguments");
        return TD_ERR_INVALID_ARG;
    }
    if (count == 0) return TD_OK;
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, count, value);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? TD_OK : cuda_fail("td_peer_wait", e);
}
