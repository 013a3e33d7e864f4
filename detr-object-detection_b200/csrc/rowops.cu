// Row-wise helper kernels of the transformer blocks (HBM-bound, coalesced 16-byte accesses):
//   colsum_bf16   bias gradient of nn.Linear: db[n] = sum_m g[m][n]  (detr/model.py:312-314,354,405-411 backward).
//                 ATen's generic reduce_kernel needs ~27 us for a 6800 x 2048 bf16 matrix; this is one pass at
//                 HBM speed: grid = column tiles x row chunks, fp32 partials, fixed-order second stage (deterministic).
#include "common.cuh"
#include <cuda_bf16.h>

namespace detr {

constexpr int kColsPerCta = 256;   // 32 lanes x 8 bf16 (16 bytes)
constexpr int kCsThreads = 256;    // 8 warps: each warp takes every 8th row of the chunk

__global__ void __launch_bounds__(kCsThreads) colsum_partial_kernel(const __nv_bfloat16* __restrict__ g, int64_t ld, int M, int N,
                                                                      int rows_per_cta, float* __restrict__ partial) {
    __shared__ float red[kCsThreads / 32][kColsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * kColsPerCta + lane * 8;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c0 < N) {   // N is a multiple of 8
        for (int r = r0 + warp; r < r1; r += kCsThreads / 32) {
            const uint4 v = *reinterpret_cast<const uint4*>(g + (int64_t)r * ld + c0);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h[e]);
                acc[2 * e] += f.x; acc[2 * e + 1] += f.y;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[warp][lane * 8 + e] = acc[e];
    __syncthreads();
    const int c = threadIdx.x;
    if (blockIdx.x * kColsPerCta + c < N) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kCsThreads / 32; ++w) s += red[w][c];
        partial[(int64_t)blockIdx.y * N + blockIdx.x * kColsPerCta + c] = s;
    }
}

__global__ void colsum_final_kernel(const float* __restrict__ partial, int chunks, int N, float* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int k = 0; k < chunks; ++k) s += partial[(int64_t)k * N + n];
    out[n] = s;
}

}  // namespace detr

using namespace detr;

extern "C" int detr_colsum_chunks(int M, int N) {
    const int col_tiles = (N + kColsPerCta - 1) / kColsPerCta;
    int chunks = (4 * 148 + col_tiles - 1) / col_tiles;          // ~4 CTAs per SM in flight
    const int max_chunks = (M + 63) / 64;                        // at least 64 rows per CTA
    if (chunks > max_chunks) chunks = max_chunks;
    return chunks < 1 ? 1 : chunks;
}

extern "C" int detr_colsum_bf16(const void* g, int64_t ld, int M, int N, float* partial, float* out, void* stream) {
    DETR_CHECK_ARG(M >= 1 && N >= 8 && (N % 8) == 0 && (ld % 8) == 0 && ((uintptr_t)g % 16) == 0,
                   "colsum: need N %% 8 == 0, ld %% 8 == 0 and a 16-byte aligned matrix (M=%d N=%d ld=%lld)", M, N, (long long)ld);
    const int chunks = detr_colsum_chunks(M, N);
    const int rows_per_cta = (M + chunks - 1) / chunks;
    dim3 grid((N + kColsPerCta - 1) / kColsPerCta, chunks);
    colsum_partial_kernel<<<grid, kCsThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(g), ld, M, N, rows_per_cta, partial);
    DETR_CHECK_LAUNCH("colsum_partial");
    colsum_final_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(partial, chunks, N, out);
    DETR_CHECK_LAUNCH("colsum_final");
    return 0;
}
