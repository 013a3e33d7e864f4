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

}  // namespace detr
