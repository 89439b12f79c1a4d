// Shared helpers for libsea_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <utility>

#include "../../include/sea_b200.h"

namespace sea {

void set_error(const char* fmt, ...);

#define SEA_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            sea::set_error(__VA_ARGS__);         \
            return SEA_ERR_INVALID;              \
        }                                        \
    } while (0)

#define SEA_CHECK_LAUNCH(name)                                                         \
    do {                                                                               \
        cudaError_t e_ = cudaGetLastError();                                           \
        if (e_ != cudaSuccess) {                                                       \
            sea::set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));     \
            return SEA_ERR_CUDA;                                                       \
        }                                                                              \
    } while (0)

#define SEA_CUDA_TRY(expr, name)                                                       \
    do {                                                                               \
        cudaError_t e_ = (expr);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            sea::set_error("%s: %s", name, cudaGetErrorString(e_));                    \
            return SEA_ERR_CUDA;                                                       \
        }                                                                              \
    } while (0)

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

template <typename T>
__device__ __forceinline__ float to_f32(T x);
template <>
__device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half x) { return __half2float(x); }

template <typename T>
__device__ __forceinline__ T from_f32(float x);
template <>
__device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <>
__device__ __forceinline__ __half from_f32<__half>(float x) { return __float2half_rn(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
// inclusive scan across the warp
__device__ __forceinline__ int warp_scan_incl_i(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

template <typename IdxT>
__device__ __forceinline__ int64_t ld_idx(const void* p, int64_t i) {
    return (int64_t) reinterpret_cast<const IdxT*>(p)[i];
}

// Dispatch a lambda on the activation dtype.
#define SEA_DISPATCH_DTYPE(dtype, T, ...)                                   \
    do {                                                                    \
        if ((dtype) == SEA_DTYPE_F32) { using T = float; __VA_ARGS__; }     \
        else if ((dtype) == SEA_DTYPE_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
        else if ((dtype) == SEA_DTYPE_F16) { using T = __half; __VA_ARGS__; } \
        else { sea::set_error("unsupported dtype %d", (int)(dtype)); return SEA_ERR_INVALID; } \
    } while (0)

#define SEA_DISPATCH_IDX(idx64, I, ...)                      \
    do {                                                     \
        if (idx64) { using I = int64_t; __VA_ARGS__; }       \
        else { using I = int32_t; __VA_ARGS__; }             \
    } while (0)

inline int cdiv(int64_t a, int64_t b) { return (int) ((a + b - 1) / b); }


// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------------
// Kernels of the layer's main chain are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel may be
// scheduled while its predecessor in the stream still runs (its CTAs become resident as the predecessor's drain), and blocks in
// pdl_wait() until the predecessor grid has completed and its memory is visible.  Rules kept by every such kernel: (1)
// pdl_wait() precedes every access to activations, reads AND writes (the allocator may recycle a buffer the predecessor still
// reads); only parameters (weights, biases) and shape arithmetic may come before it; (2) every kernel calls pdl_wait() in at
// least the threads that touch global memory, so completion stays transitive along the chain; (3) pdl_launch_dependents() at
// the top lets the successor start early.  SEA_NO_PDL=1 launches without the attribute (A/B timing).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// launch_pdl_if(allow, ...): allow = false gives a plain, fully serialised launch.  Needed when the predecessor in the stream wrote
// something the kernel reads BEFORE its pdl_wait() -- i.e. a weight-packing kernel launched just before a kernel that prefetches
// its packed weights during the set-up.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_if(bool allow, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    static const bool no_pdl = getenv("SEA_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (no_pdl || !allow) ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    static const bool no_pdl = getenv("SEA_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace sea
