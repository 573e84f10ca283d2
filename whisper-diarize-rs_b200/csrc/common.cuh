// common.cuh — error plumbing, launch accounting and small device helpers shared by every kernel file.
#pragma once
#include <utility>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <mutex>
#include "../../include/wdr.h"

namespace wdr {

// thread-local last-error text (wdr_last_error)
void set_error(const char* fmt, ...);
void clear_error();
void log_msg(int level, const char* fmt, ...);

extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Returns WDR_OK if a usable device exists (and makes `device` current), WDR_ERR_NO_DEVICE otherwise.
int ensure_device(int device);

#define WDR_CUDA_TRY(expr)                                                                          \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::wdr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (_e == cudaErrorMemoryAllocation) ? WDR_ERR_OOM : WDR_ERR_CUDA;                  \
        }                                                                                           \
    } while (0)

#define WDR_LAUNCH_CHECK()                                                                                 \
    do {                                                                                                   \
        ::wdr::count_launch();                                                                             \
        cudaError_t _e = cudaGetLastError();                                                               \
        if (_e != cudaSuccess) {                                                                           \
            ::wdr::set_error("%s:%d: kernel launch failed -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return WDR_ERR_CUDA;                                                                           \
        }                                                                                                  \
    } while (0)

#define WDR_REQUIRE(cond, msg)                   \
    do {                                         \
        if (!(cond)) {                           \
            ::wdr::set_error("%s: %s", __func__, msg); \
            return WDR_ERR_INVALID;              \
        }                                        \
    } while (0)

// Block cache behind DevBuf (core.cu): cudaMalloc / cudaFree of the host-pointer entry points' temporaries showed erratic
// 10-400 ms stalls per call on the B200 boxes, so released blocks are kept (per device, bucketed by size) and handed out again.
void* devbuf_acquire(size_t bytes);   // nullptr on allocation failure (last error set by cudaMalloc)
void devbuf_release(void* p);         // synchronises the device first (what cudaFree did implicitly), then caches the block
void devbuf_trim();                   // frees every cached block (called when a context / model is destroyed)

// RAII device buffer for the host-pointer entry points.
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) devbuf_release(p); }
    cudaError_t alloc(size_t count) {
        n = count;
        if (p) { devbuf_release(p); p = nullptr; }
        p = reinterpret_cast<T*>(devbuf_acquire((count ? count : 1) * sizeof(T)));
        return p ? cudaSuccess : cudaErrorMemoryAllocation;
    }
};

// Monotone float <-> uint key (for atomicMax over floats of either sign).
__device__ __forceinline__ unsigned float_to_key(float f) {
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- programmatic dependent launch (PDL) ----
// A kernel launched with the programmatic-stream-serialization attribute may start while its predecessor still runs:
// pdl_launch_dependents() (first statement) lets the successor's CTAs be scheduled once every CTA of this grid has started;
// pdl_wait() blocks until the predecessor grid has completed and its memory is visible.  Everything before pdl_wait() must
// neither read the predecessor's output nor write anything it may still read.  Without the attribute both are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

// Function attributes (cudaFuncSetAttribute) and device properties belong to ONE device: a process that opens contexts on several
// devices (wdr_context_params.gpu_device) must set them on each.  f() runs once per (call site, current device).
struct DeviceOnce {
    std::mutex mu;
    uint64_t done = 0;
};
template <typename F>
inline cudaError_t per_device_once(DeviceOnce& o, F f) {
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    std::lock_guard<std::mutex> lk(o.mu);
    if (o.done & bit) return cudaSuccess;
    const cudaError_t e = f();
    if (e == cudaSuccess) o.done |= bit;
    return e;
}

}  // namespace wdr
