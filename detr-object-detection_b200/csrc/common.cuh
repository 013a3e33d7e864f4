// Shared helpers for libdetr_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/detr_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libdetr_b200 is written for sm_100a only"
#endif

namespace detr {

void set_error(const char* fmt, ...);

#define DETR_CHECK_ARG(cond, ...)      \
    do {                               \
        if (!(cond)) {                 \
            detr::set_error(__VA_ARGS__); \
            return 1;                  \
        }                              \
    } while (0)

#define DETR_CHECK_LAUNCH(name)                                                          \
    do {                                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            detr::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));     \
            return 2;                                                                    \
        }                                                                                \
    } while (0)

constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------
// A kernel launched through launch_pdl may be scheduled while its predecessor in the stream is still running (its CTAs
// become resident as SMs free up, so launch latency and prologue overlap the predecessor's tail); it must execute
// pdl_wait() before it touches anything the predecessor wrote -- the wait returns once the predecessor grid has completed
// and its writes are visible.  pdl_trigger() in the predecessor allows the scheduling early; both are no-ops for kernels
// launched the ordinary way.  DETR_B200_NO_PDL=1 turns launch_pdl into an ordinary launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_if(bool on, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (on && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace detr
