// Multi-tensor "scale output channel and cast" for the frozen-BatchNorm fold of the backbone's callers
// (detr/model.py:427-438: torchvision ResNet with FrozenBatchNorm2d).  conv(x, W) * s + t == conv(x, W * s) + t, so the
// harness feeds cuDNN bf16 weights W * s.  Doing that per convolution costs 3 launches forward (mul, cast, layout copy)
// and 3 backward for each of the 53 convolutions, every step; here ONE launch folds every weight of the network
// (fp32 any-layout -> bf16 OHWI) and ONE launch turns the 53 bf16 weight gradients back into fp32 parameter gradients
// (dW = dW' * s).  Up to 64 tensors per launch, described by a table passed in the kernel parameters.
#include "common.cuh"
#include <cuda_bf16.h>

namespace detr {

template <typename T> __device__ __forceinline__ float fold_ld(const void* p, int64_t i);
template <> __device__ __forceinline__ float fold_ld<float>(const void* p, int64_t i) { return reinterpret_cast<const float*>(p)[i]; }
template <> __device__ __forceinline__ float fold_ld<__nv_bfloat16>(const void* p, int64_t i) { return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]); }
template <typename T> __device__ __forceinline__ void fold_st(void* p, int64_t i, float v);
template <> __device__ __forceinline__ void fold_st<float>(void* p, int64_t i, float v) { reinterpret_cast<float*>(p)[i] = v; }
template <> __device__ __forceinline__ void fold_st<__nv_bfloat16>(void* p, int64_t i, float v) { reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16(v); }

// blockIdx.y = tensor; the CTAs of a row grid-stride over the tensor in (o, hw, i) order (i fastest: contiguous for
// channels_last tensors on both sides)
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) scale_cast_multi_kernel(const DetrFoldTable t) {
    const int k = blockIdx.y;
    const int O = t.O[k], I = t.I[k], HW = t.HW[k];
    const int64_t n = (int64_t)O * I * HW;
    const void* src = t.src[k];
    void* dst = t.dst[k];
    const float* scale = t.scale[k];
    const int s_o = t.src_stride[k][0], s_i = t.src_stride[k][1], s_hw = t.src_stride[k][2];
    const int d_o = t.dst_stride[k][0], d_i = t.dst_stride[k][1], d_hw = t.dst_stride[k][2];
    const int row = I * HW;   // elements per output channel
    if (s_i == 1 && d_i == 1 && s_hw == I && d_hw == I && s_o == row && d_o == row && (row & 3) == 0 &&
        ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0) {
        // both sides dense in (o, hw, i) order (channels_last weights): 4 elements per thread, one division per group
        for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n; q += (int64_t)gridDim.x * blockDim.x) {
            const int64_t idx = q * 4;
            const float sc = scale[(int)(idx / row)];
            float v[4];
            if (sizeof(TIn) == 4) {
                const float4 a = reinterpret_cast<const float4*>(src)[q];
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            } else {
                const uint2 a = reinterpret_cast<const uint2*>(src)[q];
                const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a.x));
                const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a.y));
                v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
            }
            if (sizeof(TOut) == 4) {
                reinterpret_cast<float4*>(dst)[q] = make_float4(v[0] * sc, v[1] * sc, v[2] * sc, v[3] * sc);
            } else {
                uint2 o;
                *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(v[0] * sc, v[1] * sc);
                *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(v[2] * sc, v[3] * sc);
                reinterpret_cast<uint2*>(dst)[q] = o;
            }
        }
        return;
    }
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % I);
        const int64_t r = idx / I;
        const int hw = (int)(r % HW), o = (int)(r / HW);
        const float v = fold_ld<TIn>(src, (int64_t)o * s_o + (int64_t)i * s_i + (int64_t)hw * s_hw) * scale[o];
        fold_st<TOut>(dst, (int64_t)o * d_o + (int64_t)i * d_i + (int64_t)hw * d_hw, v);
    }
}

}  // namespace detr

using namespace detr;

extern "C" int detr_scale_cast_multi(const DetrFoldTable* table, int in_dtype, int out_dtype, void* stream) {
    DETR_CHECK_ARG(table != nullptr && table->n >= 1 && table->n <= DETR_FOLD_MAX_TENSORS, "scale_cast_multi: 1..%d tensors per call", DETR_FOLD_MAX_TENSORS);
    DETR_CHECK_ARG((in_dtype == 0 || in_dtype == 1) && (out_dtype == 0 || out_dtype == 1), "scale_cast_multi: dtype codes are 0 (float32) / 1 (bfloat16)");
    int64_t biggest = 0;
    for (int k = 0; k < table->n; ++k) {
        DETR_CHECK_ARG(table->src[k] && table->dst[k] && table->scale[k] && table->O[k] >= 1 && table->I[k] >= 1 && table->HW[k] >= 1,
                       "scale_cast_multi: bad entry %d", k);
        const int64_t n = (int64_t)table->O[k] * table->I[k] * table->HW[k];
        if (n > biggest) biggest = n;
    }
    int gx = (int)((biggest + 256 * 8 - 1) / (256 * 8));   // ~8 elements per thread on the largest tensor
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    dim3 grid(gx, table->n);
    cudaStream_t st = (cudaStream_t)stream;
    if (in_dtype == 0 && out_dtype == 1) scale_cast_multi_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(*table);
    else if (in_dtype == 1 && out_dtype == 0) scale_cast_multi_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(*table);
    else if (in_dtype == 0) scale_cast_multi_kernel<float, float><<<grid, 256, 0, st>>>(*table);
    else scale_cast_multi_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(*table);
    DETR_CHECK_LAUNCH("scale_cast_multi");
    return 0;
}
