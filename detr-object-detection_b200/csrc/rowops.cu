// Row-wise helper kernels of the transformer blocks (HBM-bound, coalesced 16-byte accesses):
//   colsum_bf16   bias gradient of nn.Linear: db[n] = sum_m g[m][n]  (detr/model.py:312-314,354,405-411 backward).
//                 ATen's generic reduce_kernel needs ~27 us for a 6800 x 2048 bf16 matrix; this is one pass at
//                 HBM speed: grid = column tiles x (<= 64) row chunks, fp32 partials; the last CTA of a column tile folds
//                 them in a fixed order with its 8 warps in parallel (deterministic, one launch).  (A serial fold of
//                 hundreds of partials by one thread per column cost 25-40 us per call, measured; LayerNorm's 296 CTA
//                 partials go through fold_partials_kernel instead.)
#include "common.cuh"
#include <cuda_bf16.h>
#include "tc.cuh"
#include "rowmath.cuh"

namespace detr {

constexpr int kColsPerCta = 256;   // 32 lanes x 8 bf16 (16 bytes)
constexpr int kCsThreads = 256;    // 8 warps: each warp takes every 8th row of the chunk, 4 rows in flight

// out[c] = sum_k partial[k][c] (k < n_partials, row stride N), columns c < split go to out0, the rest to out1.
// CTA = 32 columns; warp w adds partials w, w+8, ... (independent coalesced loads), then the 8 warps are combined.
constexpr int kFoldThreads = 1024;
__global__ void __launch_bounds__(kFoldThreads) fold_partials_kernel(const float* __restrict__ partial, int n_partials, int N,
                                                                     float* __restrict__ out0, float* __restrict__ out1, int split,
                                                                     float* __restrict__ out2 = nullptr) {
    __shared__ float red[kFoldThreads / 32][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (c < N) {
        constexpr int W = kFoldThreads / 32;
        int k = warp;
        for (; k + 3 * W < n_partials; k += 4 * W) {   // four independent loads in flight per thread
            s0 += __ldcg(&partial[(int64_t)k * N + c]);
            s1 += __ldcg(&partial[(int64_t)(k + W) * N + c]);
            s2 += __ldcg(&partial[(int64_t)(k + 2 * W) * N + c]);
            s3 += __ldcg(&partial[(int64_t)(k + 3 * W) * N + c]);
        }
        for (; k < n_partials; k += W) s0 += __ldcg(&partial[(int64_t)k * N + c]);
    }
    red[warp][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (warp == 0 && c < N) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kFoldThreads / 32; ++w) s += red[w][lane];
        if (c < split) out0[c] = s; else if (c < 2 * split || out2 == nullptr) out1[c - split] = s; else out2[c - 2 * split] = s;
    }
}

// Tail of the column-sum kernels: the last CTA of a column tile (grid.y <= 64 row chunks) folds the chunk partials in a
// fixed order -- 8 warps take every 8th chunk with independent 32-byte loads, then combine through shared memory -- so
// the bias gradient needs no second launch.  counter: one uint32 per column tile, zero on entry and on exit.
__device__ __forceinline__ void colsum_fold_tail(const float* __restrict__ partial, int N, float* __restrict__ out,
                                                 unsigned* __restrict__ counter, float (*red)[kColsPerCta]) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.y - 1;
    __syncthreads();
    if (!is_last) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * kColsPerCta + lane * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c0 < N) {
        constexpr int W = kCsThreads / 32;
        for (int k = warp; k < (int)gridDim.y; k += 2 * W) {
            const float* r0 = partial + (int64_t)k * N + c0;
            const bool two = k + W < (int)gridDim.y;
            const float* r1 = two ? r0 + (int64_t)W * N : r0;
            const float4 a0 = __ldcg(reinterpret_cast<const float4*>(r0)), a1 = __ldcg(reinterpret_cast<const float4*>(r0) + 1);
            const float4 b0 = __ldcg(reinterpret_cast<const float4*>(r1)), b1 = __ldcg(reinterpret_cast<const float4*>(r1) + 1);
            acc[0] += a0.x; acc[1] += a0.y; acc[2] += a0.z; acc[3] += a0.w; acc[4] += a1.x; acc[5] += a1.y; acc[6] += a1.z; acc[7] += a1.w;
            if (two) { acc[0] += b0.x; acc[1] += b0.y; acc[2] += b0.z; acc[3] += b0.w; acc[4] += b1.x; acc[5] += b1.y; acc[6] += b1.z; acc[7] += b1.w; }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[warp][lane * 8 + e] = acc[e];
    __syncthreads();
    const int c = threadIdx.x, col = blockIdx.x * kColsPerCta + c;
    if (c < kColsPerCta && col < N) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kCsThreads / 32; ++w) s += red[w][c];
        out[col] = s;
    }
    if (threadIdx.x == 0) *counter = 0;
}

__global__ void __launch_bounds__(kCsThreads) colsum_kernel(const __nv_bfloat16* __restrict__ g, int64_t ld, int M, int N, int rows_per_cta,
                                                              float* __restrict__ partial, float* __restrict__ out, unsigned* __restrict__ counters) {
    __shared__ float red[kCsThreads / 32][kColsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * kColsPerCta + lane * 8;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c0 < N) {   // N is a multiple of 8
        constexpr int W = kCsThreads / 32;
        for (int r = r0 + warp; r < r1; r += 4 * W) {    // four independent 16-byte loads in flight per lane
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                v[u] = (r + u * W < r1) ? *reinterpret_cast<const uint4*>(g + (int64_t)(r + u * W) * ld + c0) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(h[e]);
                    acc[2 * e] += f.x; acc[2 * e + 1] += f.y;
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[warp][lane * 8 + e] = acc[e];
    __syncthreads();
    const int c = threadIdx.x, col = blockIdx.x * kColsPerCta + c;
    if (c < kColsPerCta && col < N) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kCsThreads / 32; ++w) s += red[w][c];
        partial[(int64_t)blockIdx.y * N + col] = s;
    }
    colsum_fold_tail(partial, N, out, counters + blockIdx.x, red);
}

}  // namespace detr

using namespace detr;

extern "C" int detr_colsum_chunks(int M, int N) {
    const int col_tiles = (N + kColsPerCta - 1) / kColsPerCta;
    int chunks = (4 * 148 + col_tiles - 1) / col_tiles;          // ~4 CTAs of 8 warps per SM
    const int max_chunks = (M + 31) / 32;                        // at least 32 rows (one 4-row batch per warp) per CTA
    if (chunks > 64) chunks = 64;                                // bounds the in-kernel fold: 8 partials per warp of the last CTA
    if (chunks > max_chunks) chunks = max_chunks;
    return chunks < 1 ? 1 : chunks;
}

extern "C" int detr_colsum_bf16(const void* g, int64_t ld, int M, int N, float* partial, float* out, uint32_t* counters, void* stream) {
    DETR_CHECK_ARG(M >= 1 && N >= 8 && (N % 8) == 0 && (ld % 8) == 0 && ((uintptr_t)g % 16) == 0,
                   "colsum: need N %% 8 == 0, ld %% 8 == 0 and a 16-byte aligned matrix (M=%d N=%d ld=%lld)", M, N, (long long)ld);
    DETR_CHECK_ARG(counters != nullptr && N <= 64 * kColsPerCta, "colsum: counters required (64 zeroed uint32), N <= %d", 64 * kColsPerCta);
    const int chunks = detr_colsum_chunks(M, N);
    const int rows_per_cta = (M + chunks - 1) / chunks;
    dim3 grid((N + kColsPerCta - 1) / kColsPerCta, chunks);
    colsum_kernel<<<grid, kCsThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(g), ld, M, N, rows_per_cta, partial, out, counters);
    DETR_CHECK_LAUNCH("colsum");
    return 0;
}

// =========================================================================================================
// Fused pre-LN prologue (detr/model.py:221-222, 173-174, 177-178, 224, 182, 209, 148):
//   y  = LayerNorm(x) * gamma + beta                      (value input of the attention / input of the FFN)
//   y2 = y + addend                                       (query/key input: + positional or query embedding)
// one pass over x, both outputs written in the GEMM's input dtype (no fp32 round trip, no separate add / cast
// kernels).  One warp per row, the row lives in registers, two-pass variance.  Backward: dx and per-CTA partial
// dgamma / dbeta, folded by fold_partials_kernel (second launch, deterministic).
// =========================================================================================================
namespace detr {

constexpr int kLnThreads = 256;   // 8 warps, one row per warp at a time
constexpr int kLnMaxPerLane = 32;   // C <= 1024

template <typename T> __device__ __forceinline__ float ld1(const T* p);
template <> __device__ __forceinline__ float ld1<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st1(T* p, float v);
template <> __device__ __forceinline__ void st1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

// lane owns the contiguous columns [lane*K, lane*K + K), K = C/32; 8-element (16/32-byte) vector accesses when K % 8 == 0
template <typename T> __device__ __forceinline__ void load_row(const T* row, int lane, int K, float* v) {
    const T* p = row + lane * K;
    if ((K & 7) == 0) {
        _Pragma("unroll") for (int i = 0; i < K; i += 8) {
            if (sizeof(T) == 2) {
                const uint4 u = *reinterpret_cast<const uint4*>(p + i);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h[e]); v[i + 2 * e] = f.x; v[i + 2 * e + 1] = f.y; }
            } else {
                const float4 a = *reinterpret_cast<const float4*>(p + i), b = *reinterpret_cast<const float4*>(p + i + 4);
                v[i] = a.x; v[i + 1] = a.y; v[i + 2] = a.z; v[i + 3] = a.w; v[i + 4] = b.x; v[i + 5] = b.y; v[i + 6] = b.z; v[i + 7] = b.w;
            }
        }
    } else {
        _Pragma("unroll") for (int i = 0; i < K; ++i) v[i] = ld1<T>(p + i);
    }
}
template <typename T> __device__ __forceinline__ void store_row(T* row, int lane, int K, const float* v) {
    T* p = row + lane * K;
    if ((K & 7) == 0) {
        _Pragma("unroll") for (int i = 0; i < K; i += 8) {
            if (sizeof(T) == 2) {
                uint4 u;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[i + 2 * e], v[i + 2 * e + 1]);
                *reinterpret_cast<uint4*>(p + i) = u;
            } else {
                *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                *reinterpret_cast<float4*>(p + i + 4) = make_float4(v[i + 4], v[i + 5], v[i + 6], v[i + 7]);
            }
        }
    } else {
        _Pragma("unroll") for (int i = 0; i < K; ++i) st1<T>(p + i, v[i]);
    }
}

struct LnParams {
    const void* x; int64_t x_ld;
    const float* gamma; const float* beta;
    const void* addend; int64_t add_sb, add_sr; int rows_per_batch;   // addend row of (b, r) = addend + b*add_sb + r*add_sr (elements)
    void* y; void* y2;                                                 // either may be null
    float* mean; float* rstd;
    int rows, C; float eps;
    // backward
    const void* dy; const void* dy2; void* dx;
    const void* dres;   // optional gradient of the residual branch that bypasses the LayerNorm (x's dtype, contiguous rows): dx += dres
    float* partial; float* dgamma; float* dbeta; unsigned* counters;
    // optional third product of the backward pass: dx is also the gradient that reaches the PREVIOUS block's tail
    // (x = res + dropout(z), detr/model.py:223-224); its masked bf16 form dz = mask(dx) / (1 - p) and the bias gradient
    // column sums of dz are produced here, so that the tail's backward needs no pass of its own
    void* dz; float* dbias; uint32_t thr4; float scale; uint64_t seed; const uint64_t* seed_ptr; int has_dz;
    int early_trigger;   // small launch (decoder rows): programmatic dependents may be scheduled at once
};

// KT > 0: compile-time columns per lane (rows stay in registers); KT == 0: run-time (local-memory arrays)
template <typename TX, typename TO, int KT>
__global__ void __launch_bounds__(kLnThreads) ln_fwd_kernel(const LnParams p) {
    using TA = float;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = KT ? KT : p.C / 32;
    constexpr int KA = KT ? KT : kLnMaxPerLane;
    float v[KA], g[KA], b[KA];
    load_row<float>(p.gamma, lane, K, g);
    load_row<float>(p.beta, lane, K, b);
    for (int r = blockIdx.x * (kLnThreads / 32) + warp; r < p.rows; r += gridDim.x * (kLnThreads / 32)) {
        load_row<TX>(reinterpret_cast<const TX*>(p.x) + (int64_t)r * p.x_ld, lane, K, v);
        float s = 0.f;
        _Pragma("unroll") for (int i = 0; i < K; ++i) s += v[i];
        const float mu = warp_sum(s) / (float)p.C;
        float q = 0.f;
        _Pragma("unroll") for (int i = 0; i < K; ++i) { const float d = v[i] - mu; q += d * d; }
        const float rs = rsqrtf(warp_sum(q) / (float)p.C + p.eps);
        if (lane == 0) { p.mean[r] = mu; p.rstd[r] = rs; }
        _Pragma("unroll") for (int i = 0; i < K; ++i) v[i] = (v[i] - mu) * rs * g[i] + b[i];
        if (p.y) store_row<TO>(reinterpret_cast<TO*>(p.y) + (int64_t)r * p.C, lane, K, v);
        if (p.y2) {
            float a[KA];
            const int bb = r / p.rows_per_batch, rr = r - bb * p.rows_per_batch;
            load_row<TA>(reinterpret_cast<const TA*>(p.addend) + bb * p.add_sb + rr * p.add_sr, lane, K, a);
            _Pragma("unroll") for (int i = 0; i < K; ++i) v[i] += a[i];
            store_row<TO>(reinterpret_cast<TO*>(p.y2) + (int64_t)r * p.C, lane, K, v);
        }
    }
}

// dy / dy2: gradients w.r.t. y / y2 (type TG, row stride C, either may be null); dx in TX
template <typename TX, typename TG, int KT>
__global__ void __launch_bounds__(kLnThreads) ln_bwd_kernel(const LnParams p) {
    extern __shared__ float sm[];   // [warps][2][C]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = KT ? KT : p.C / 32, C = p.C;
    constexpr int KA = KT ? KT : kLnMaxPerLane;
    float g[KA], dg[KA], db[KA], dzs[KA];
    load_row<float>(p.gamma, lane, K, g);           // a parameter: not written by the kernel in front of this one
    pdl_wait();
    if (p.early_trigger) pdl_trigger();
    _Pragma("unroll") for (int i = 0; i < K; ++i) { dg[i] = 0.f; db[i] = 0.f; dzs[i] = 0.f; }
    const bool has_dz = KT == 8 && p.has_dz;   // (the 8-columns-per-lane layout is exactly one dropout chunk per lane)
    const uint32_t zkey = (has_dz && p.thr4) ? ew_key(p.seed, p.seed_ptr) : 0u;
    for (int r = blockIdx.x * (kLnThreads / 32) + warp; r < p.rows; r += gridDim.x * (kLnThreads / 32)) {
        float x[KA], d[KA];
        load_row<TX>(reinterpret_cast<const TX*>(p.x) + (int64_t)r * p.x_ld, lane, K, x);
        if (p.dy) load_row<TG>(reinterpret_cast<const TG*>(p.dy) + (int64_t)r * C, lane, K, d);
        else _Pragma("unroll") for (int i = 0; i < K; ++i) d[i] = 0.f;
        if (p.dy2) {
            float d2[KA];
            load_row<TG>(reinterpret_cast<const TG*>(p.dy2) + (int64_t)r * C, lane, K, d2);
            _Pragma("unroll") for (int i = 0; i < K; ++i) d[i] += d2[i];
        }
        const float mu = p.mean[r], rs = p.rstd[r];
        float c1 = 0.f, c2 = 0.f;
        _Pragma("unroll") for (int i = 0; i < K; ++i) {
            x[i] = (x[i] - mu) * rs;            // xhat
            dg[i] += d[i] * x[i];
            db[i] += d[i];
            d[i] *= g[i];                       // wdy
            c1 += d[i] * x[i];
            c2 += d[i];
        }
        c1 = warp_sum(c1) / (float)C;
        c2 = warp_sum(c2) / (float)C;
        _Pragma("unroll") for (int i = 0; i < K; ++i) d[i] = rs * (d[i] - c2 - x[i] * c1);
        if (p.dres) {   // x also feeds the block's residual add: its gradient arrives here instead of in a separate add kernel
            float e[KA];
            load_row<TX>(reinterpret_cast<const TX*>(p.dres) + (int64_t)r * C, lane, K, e);
            _Pragma("unroll") for (int i = 0; i < K; ++i) d[i] += e[i];
        }
        store_row<TX>(reinterpret_cast<TX*>(p.dx) + (int64_t)r * C, lane, K, d);
        if (has_dz) {
            // the same mask as the tail's forward epilogue: chunk index r * C/8 + lane, 8 columns per chunk
            if (p.thr4) {
                bool keep[8];
                ew_keep8(zkey, (uint32_t)r * (uint32_t)(C >> 3) + (uint32_t)lane, p.thr4, keep);
                _Pragma("unroll") for (int i = 0; i < 8; ++i) d[i] = keep[i] ? d[i] * p.scale : 0.f;
            }
            uint4 u;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
            _Pragma("unroll") for (int e = 0; e < 4; ++e) {
                h[e] = __floats2bfloat162_rn(d[2 * e], d[2 * e + 1]);
                const float2 f = __bfloat1622float2(h[e]);      // the bias gradient sums what the GEMMs will see
                dzs[2 * e] += f.x; dzs[2 * e + 1] += f.y;
            }
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.dz) + (int64_t)r * C + lane * 8) = u;
        }
    }
    // ---- dgamma / dbeta (/ dbias): warps -> CTA partial; fold_partials_kernel adds the CTA partials ----
    const int NP = has_dz ? 3 : 2;
    _Pragma("unroll") for (int i = 0; i < K; ++i) {
        sm[(warp * 3 + 0) * C + lane * K + i] = dg[i]; sm[(warp * 3 + 1) * C + lane * K + i] = db[i];
        if (has_dz) sm[(warp * 3 + 2) * C + lane * K + i] = dzs[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < NP * C; c += kLnThreads) {
        const int which = c / C, col = c - which * C;
        float s = 0.f;
        for (int w = 0; w < kLnThreads / 32; ++w) s += sm[(w * 3 + which) * C + col];
        p.partial[(int64_t)blockIdx.x * NP * C + c] = s;
    }
}

static int ln_grid(int rows) {   // backward: <= 296 CTA partials of dgamma / dbeta for the fold
    int g = (rows + kLnThreads / 32 - 1) / (kLnThreads / 32);
    return g > 2 * 148 ? 2 * 148 : (g < 1 ? 1 : g);
}
static int ln_grid_fwd(int rows) {
    int g = (rows + kLnThreads / 32 - 1) / (kLnThreads / 32);
    return g > 8 * 148 ? 8 * 148 : (g < 1 ? 1 : g);
}

}  // namespace detr

extern "C" int detr_layernorm_grid(int rows) { return detr::ln_grid(rows); }

// dtype codes: 0 = float32, 1 = bfloat16
extern "C" int detr_layernorm_fwd(const void* x, int x_dtype, int64_t x_ld, const float* gamma, const float* beta,
                                  const void* addend, int add_dtype, int64_t add_sb, int64_t add_sr, int rows_per_batch,
                                  void* y, void* y2, int out_dtype, float* mean, float* rstd, int rows, int C, float eps, void* stream) {
    DETR_CHECK_ARG(rows >= 1 && C >= 32 && C % 32 == 0 && C <= 32 * kLnMaxPerLane, "layernorm: C=%d must be a multiple of 32, <= %d", C, 32 * kLnMaxPerLane);
    DETR_CHECK_ARG((y2 == nullptr) == (addend == nullptr) || y2 == nullptr, "layernorm: y2 needs an addend");
    DETR_CHECK_ARG(x_ld % 8 == 0 && ((uintptr_t)x % 16) == 0 && (!addend || (((uintptr_t)addend % 16) == 0 && add_sb % 8 == 0 && add_sr % 8 == 0)),
                   "layernorm: rows must be 16-byte aligned");
    LnParams p{};
    p.x = x; p.x_ld = x_ld; p.gamma = gamma; p.beta = beta; p.addend = addend; p.add_sb = add_sb; p.add_sr = add_sr;
    p.rows_per_batch = rows_per_batch > 0 ? rows_per_batch : rows; p.y = y; p.y2 = y2; p.mean = mean; p.rstd = rstd; p.rows = rows; p.C = C; p.eps = eps;
    const int grid = ln_grid_fwd(rows);
    cudaStream_t st = (cudaStream_t)stream;
    DETR_CHECK_ARG(!addend || add_dtype == 0, "layernorm: the addend must be float32");
#define LN_FWD(TX, TO)                                                            \
    do {                                                                          \
        if (C == 256) ln_fwd_kernel<TX, TO, 8><<<grid, kLnThreads, 0, st>>>(p);   \
        else ln_fwd_kernel<TX, TO, 0><<<grid, kLnThreads, 0, st>>>(p);            \
    } while (0)
    if (x_dtype == 0 && out_dtype == 0) LN_FWD(float, float);
    else if (x_dtype == 0) LN_FWD(float, __nv_bfloat16);
    else if (out_dtype == 0) LN_FWD(__nv_bfloat16, float);
    else LN_FWD(__nv_bfloat16, __nv_bfloat16);
#undef LN_FWD
    DETR_CHECK_LAUNCH("layernorm_fwd");
    return 0;
}

static int layernorm_bwd_impl(const void* dy, const void* dy2, int g_dtype, const void* dres, const void* x, int x_dtype, int64_t x_ld,
                              const float* gamma, const float* mean, const float* rstd, void* dx, float* partial,
                              float* dgamma, float* dbeta, uint32_t* counters, int rows, int C, void* stream,
                              void* dz, float* dbias, float dropout_p, uint64_t seed, const uint64_t* seed_ptr);

extern "C" int detr_layernorm_bwd(const void* dy, const void* dy2, int g_dtype, const void* dres, const void* x, int x_dtype, int64_t x_ld,
                                  const float* gamma, const float* mean, const float* rstd, void* dx, float* partial,
                                  float* dgamma, float* dbeta, uint32_t* counters, int rows, int C, void* stream) {
    return layernorm_bwd_impl(dy, dy2, g_dtype, dres, x, x_dtype, x_ld, gamma, mean, rstd, dx, partial, dgamma, dbeta, counters, rows, C, stream,
                              nullptr, nullptr, 0.f, 0, nullptr);
}

extern "C" int detr_layernorm_bwd_tail(const void* dy, const void* dy2, int g_dtype, const void* dres, const void* x, int x_dtype, int64_t x_ld,
                                       const float* gamma, const float* mean, const float* rstd, void* dx, float* partial,
                                       float* dgamma, float* dbeta, uint32_t* counters, int rows, int C, void* dz, float* dbias,
                                       float dropout_p, uint64_t seed, const uint64_t* seed_ptr, void* stream) {
    DETR_CHECK_ARG(C == 256 && dz != nullptr && dbias != nullptr && ((uintptr_t)dz % 16) == 0, "layernorm_bwd_tail: C must be 256, dz / dbias required");
    DETR_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "layernorm_bwd_tail: dropout_p must be in [0,1)");
    return layernorm_bwd_impl(dy, dy2, g_dtype, dres, x, x_dtype, x_ld, gamma, mean, rstd, dx, partial, dgamma, dbeta, counters, rows, C, stream,
                              dz, dbias, dropout_p, seed, seed_ptr);
}

static int layernorm_bwd_impl(const void* dy, const void* dy2, int g_dtype, const void* dres, const void* x, int x_dtype, int64_t x_ld,
                              const float* gamma, const float* mean, const float* rstd, void* dx, float* partial,
                              float* dgamma, float* dbeta, uint32_t* counters, int rows, int C, void* stream,
                              void* dz, float* dbias, float dropout_p, uint64_t seed, const uint64_t* seed_ptr) {
    DETR_CHECK_ARG(rows >= 1 && C >= 32 && C % 32 == 0 && C <= 32 * kLnMaxPerLane, "layernorm_bwd: bad C=%d", C);
    DETR_CHECK_ARG(dy != nullptr || dy2 != nullptr, "layernorm_bwd: no incoming gradient");
    LnParams p{};
    p.x = x; p.x_ld = x_ld; p.gamma = gamma; p.mean = const_cast<float*>(mean); p.rstd = const_cast<float*>(rstd); p.rows = rows; p.C = C;
    DETR_CHECK_ARG(((uintptr_t)dres % 16) == 0, "layernorm_bwd: dres must be 16-byte aligned");
    p.dy = dy; p.dy2 = dy2; p.dres = dres; p.dx = dx; p.partial = partial; p.dgamma = dgamma; p.dbeta = dbeta; p.counters = counters;
    p.dz = dz; p.dbias = dbias; p.has_dz = dz != nullptr; p.seed = seed; p.seed_ptr = seed_ptr;
    {
        const uint32_t th = tc::dropout_threshold(dropout_p);
        p.thr4 = th * 0x00010001u; p.scale = (float)tc::kDropOne / ((float)tc::kDropOne - (float)th);
    }
    const int grid = ln_grid(rows);
    const size_t smem = (size_t)(kLnThreads / 32) * 3 * C * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    // decoder-sized calls sit in a chain of launch-latency-bound kernels: let them be scheduled behind the GEMM in front (and
    // the GEMM behind them be scheduled at once)
    p.early_trigger = rows <= 2400 ? 1 : 0;
#define LN_BWD(TX, TG)                                                               \
    do {                                                                             \
        if (C == 256) launch_pdl_if(p.early_trigger != 0, ln_bwd_kernel<TX, TG, 8>, dim3(grid), dim3(kLnThreads), smem, st, p);   \
        else launch_pdl_if(p.early_trigger != 0, ln_bwd_kernel<TX, TG, 0>, dim3(grid), dim3(kLnThreads), smem, st, p);            \
    } while (0)
    if (x_dtype == 0 && g_dtype == 0) LN_BWD(float, float);
    else if (x_dtype == 0) LN_BWD(float, __nv_bfloat16);
    else if (g_dtype == 0) LN_BWD(__nv_bfloat16, float);
    else LN_BWD(__nv_bfloat16, __nv_bfloat16);
#undef LN_BWD
    DETR_CHECK_LAUNCH("layernorm_bwd");
    (void)counters;
    if (dgamma == nullptr) return 0;          // the caller folds the partials itself (detr_layernorm_bwd_fold, e.g. on another stream)
    const int NP = p.has_dz ? 3 : 2;
    fold_partials_kernel<<<(NP * C + 31) / 32, kFoldThreads, 0, st>>>(partial, grid, NP * C, dgamma, dbeta, C, dbias);
    DETR_CHECK_LAUNCH("layernorm_bwd_fold");
    return 0;
}

// Second half of detr_layernorm_bwd[_tail] when it was called with dgamma == NULL: fold the per-CTA partials of `rows` rows into
// dgamma, dbeta (and dbias when the call had a tail).  Parameter gradients are not on the critical path of the backward pass,
// so the host launches this on a second stream.
extern "C" int detr_layernorm_bwd_fold(const float* partial, int rows, int C, float* dgamma, float* dbeta, float* dbias, void* stream) {
    DETR_CHECK_ARG(partial != nullptr && dgamma != nullptr && dbeta != nullptr && rows >= 1 && C >= 32, "layernorm_bwd_fold: bad arguments");
    const int NP = dbias != nullptr ? 3 : 2;
    fold_partials_kernel<<<(NP * C + 31) / 32, kFoldThreads, 0, (cudaStream_t)stream>>>(partial, ln_grid(rows), NP * C, dgamma, dbeta, C, dbias);
    DETR_CHECK_LAUNCH("layernorm_bwd_fold");
    return 0;
}

// =========================================================================================================
// Block epilogues of the pre-LN layers (detr/model.py:223-224,176-182 and the FFN 405-411), one pass each:
//   MODE 0   out = x + dropout(y)                      residual add after the attention output / second FFN projection
//   MODE 1   out = dropout(gelu_tanh(y))               between the two FFN projections
// and their backward, which also produces the bias gradient of the Linear that made y (column sums of dy, as fp32
// partials per row chunk folded by fold_partials_kernel), so ATen's dropout / masked_scale / gelu / gelu_backward /
// add kernels and one colsum pass disappear:
//   MODE 0   dy = dropout_mask(g) / (1-p)              (the residual branch's gradient is g itself)
//   MODE 1   dy = dropout_mask(g) / (1-p) * gelu'(y)
// The dropout mask is the 7-bit counter-based generator of the attention kernels (tc.cuh), indexed by the 8-element
// chunk (row * N/8 + column/8): nothing is stored for backward, p is quantised to k/32768.
// =========================================================================================================
#include "tc.cuh"
#include "rowmath.cuh"

namespace detr {

struct EwParams {
    const void* x;                 // MODE 0 forward: residual (TX), contiguous (M, N)
    const __nv_bfloat16* y;        // forward: GEMM output; backward: the same tensor (MODE 1 needs it for gelu')
    void* out;                     // forward: TX (MODE 0) / bf16 (MODE 1); backward: dy bf16
    const void* g;                 // backward: incoming gradient, TG
    float* partial;                // backward: [row chunks][N] column sums of dy
    float* db; unsigned* counters; // backward: bias gradient, per-column-tile counters (zero on entry and exit)
    int M, N, rows_per_cta;
    uint32_t thr4; float scale;    // dropout: thresh * 0x00010001 (0 = off; thresh in 1/32768), 1 / keep
    uint64_t seed; const uint64_t* seed_ptr;
};

template <typename T> __device__ __forceinline__ void ld8(const T* p, float* v);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float* v);
template <> __device__ __forceinline__ void st8<float>(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(p) = u;
}

constexpr int kEwThreads = 256;

template <int MODE, typename TX>
__global__ void __launch_bounds__(kEwThreads) epilogue_fwd_kernel(const EwParams p) {
    const int64_t n8 = (int64_t)p.M * (p.N >> 3);
    const bool drop = p.thr4 != 0;
    const uint32_t key = drop ? ew_key(p.seed, p.seed_ptr) : 0u;
    for (int64_t i = (int64_t)blockIdx.x * kEwThreads + threadIdx.x; i < n8; i += (int64_t)gridDim.x * kEwThreads) {
        float y[8], o[8];
        ld8<__nv_bfloat16>(p.y + i * 8, y);
        bool keep[8];
        if (drop) ew_keep8(key, (uint32_t)i, p.thr4, keep);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float v = y[e];
            if (MODE == 1) { float t; v = gelu_tanh_fwd(v, t); }
            if (drop) v = keep[e] ? v * p.scale : 0.f;
            o[e] = v;
        }
        if (MODE == 0) {
            float x[8];
            ld8<TX>(reinterpret_cast<const TX*>(p.x) + i * 8, x);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] += x[e];
            st8<TX>(reinterpret_cast<TX*>(p.out) + i * 8, o);
        } else {
            st8<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(p.out) + i * 8, o);
        }
    }
}

// grid = (column tiles of 256, row chunks); warp w of the CTA takes rows r0 + w, r0 + w + 8, ...; lane = 8 columns
template <int MODE, typename TG>
__global__ void __launch_bounds__(kCsThreads) epilogue_bwd_kernel(const EwParams p) {
    __shared__ float red[kCsThreads / 32][kColsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * kColsPerCta + lane * 8;
    const int r0 = blockIdx.y * p.rows_per_cta, r1 = min(p.M, r0 + p.rows_per_cta);
    const bool drop = p.thr4 != 0;
    const uint32_t key = drop ? ew_key(p.seed, p.seed_ptr) : 0u;
    const int n8 = p.N >> 3;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c0 < p.N) {
        for (int r = r0 + warp; r < r1; r += kCsThreads / 32) {
            const int64_t off = (int64_t)r * p.N + c0;
            float g[8], d[8];
            ld8<TG>(reinterpret_cast<const TG*>(p.g) + off, g);
            bool keep[8];
            if (drop) ew_keep8(key, (uint32_t)((int64_t)r * n8 + (c0 >> 3)), p.thr4, keep);
            float y[8];
            if (MODE == 1) ld8<__nv_bfloat16>(p.y + off, y);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float v = g[e];
                if (drop) v = keep[e] ? v * p.scale : 0.f;
                if (MODE == 1) v *= gelu_tanh_grad(y[e]);
                d[e] = v;
            }
            uint4 u;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                h[e] = __floats2bfloat162_rn(d[2 * e], d[2 * e + 1]);
                const float2 f = __bfloat1622float2(h[e]);      // the bias gradient sums what the GEMMs will see
                acc[2 * e] += f.x; acc[2 * e + 1] += f.y;
            }
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off) = u;
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[warp][lane * 8 + e] = acc[e];
    __syncthreads();
    const int c = threadIdx.x, col = blockIdx.x * kColsPerCta + c;
    if (c < kColsPerCta && col < p.N) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kCsThreads / 32; ++w) s += red[w][c];
        p.partial[(int64_t)blockIdx.y * p.N + col] = s;
    }
    colsum_fold_tail(p.partial, p.N, p.db, p.counters + blockIdx.x, red);
}

static int ew_fill(EwParams& p, int M, int N, float dropout_p, uint64_t seed, const uint64_t* seed_ptr, const char* who) {
    if (!(M >= 1 && N >= 8 && N % 8 == 0)) { set_error("%s: need M >= 1 and N %% 8 == 0 (M=%d N=%d)", who, M, N); return 1; }
    if (!((int64_t)M * (N / 8) < (1ll << 32))) { set_error("%s: tensor too large for the 32-bit chunk counter", who); return 1; }
    if (!(dropout_p >= 0.f && dropout_p < 1.f)) { set_error("%s: dropout_p must be in [0,1)", who); return 1; }
    const uint32_t th = tc::dropout_threshold(dropout_p);
    p.M = M; p.N = N; p.thr4 = th * 0x00010001u; p.scale = (float)tc::kDropOne / ((float)tc::kDropOne - (float)th); p.seed = seed; p.seed_ptr = seed_ptr;
    return 0;
}

}  // namespace detr

/* mode 0: out(TX) = x(TX) + dropout(y);  mode 1: out(bf16) = dropout(gelu_tanh(y)) (x unused).  x_dtype: 0 f32, 1 bf16. */
extern "C" int detr_epilogue_fwd(int mode, const void* x, int x_dtype, const void* y, void* out, int M, int N, float dropout_p,
                                 uint64_t seed, const uint64_t* seed_ptr, void* stream) {
    EwParams p{};
    if (int rc = ew_fill(p, M, N, dropout_p, seed, seed_ptr, "epilogue_fwd")) return rc;
    DETR_CHECK_ARG(mode == 0 || mode == 1, "epilogue_fwd: mode must be 0 or 1");
    DETR_CHECK_ARG(((uintptr_t)y % 16) == 0 && ((uintptr_t)out % 16) == 0 && (mode == 1 || ((uintptr_t)x % 16) == 0), "epilogue_fwd: 16-byte alignment");
    p.x = x; p.y = reinterpret_cast<const __nv_bfloat16*>(y); p.out = out;
    const int64_t n8 = (int64_t)M * (N / 8);
    int grid = (int)((n8 + kEwThreads - 1) / kEwThreads);
    if (grid > 148 * 16) grid = 148 * 16;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 1) epilogue_fwd_kernel<1, float><<<grid, kEwThreads, 0, st>>>(p);
    else if (x_dtype == 0) epilogue_fwd_kernel<0, float><<<grid, kEwThreads, 0, st>>>(p);
    else epilogue_fwd_kernel<0, __nv_bfloat16><<<grid, kEwThreads, 0, st>>>(p);
    DETR_CHECK_LAUNCH("epilogue_fwd");
    return 0;
}

extern "C" int detr_epilogue_chunks(int M, int N) { return detr_colsum_chunks(M, N); }

/* dy(bf16) = dropout_mask(g)/(1-p) [* gelu'(y) in mode 1]; db[n] = sum_m dy[m][n].  g_dtype: 0 f32, 1 bf16.
 * partial float[detr_epilogue_chunks(M,N) * N] scratch. */
extern "C" int detr_epilogue_bwd(int mode, const void* g, int g_dtype, const void* y, void* dy, float* partial, float* db,
                                 uint32_t* counters, int M, int N, float dropout_p, uint64_t seed, const uint64_t* seed_ptr, void* stream) {
    EwParams p{};
    if (int rc = ew_fill(p, M, N, dropout_p, seed, seed_ptr, "epilogue_bwd")) return rc;
    DETR_CHECK_ARG(mode == 0 || mode == 1, "epilogue_bwd: mode must be 0 or 1");
    DETR_CHECK_ARG(((uintptr_t)g % 16) == 0 && ((uintptr_t)dy % 16) == 0 && (mode == 0 || ((uintptr_t)y % 16) == 0), "epilogue_bwd: 16-byte alignment");
    DETR_CHECK_ARG(counters != nullptr && N <= 64 * kColsPerCta, "epilogue_bwd: counters required (64 zeroed uint32), N <= %d", 64 * kColsPerCta);
    p.g = g; p.y = reinterpret_cast<const __nv_bfloat16*>(y); p.out = dy; p.partial = partial; p.db = db; p.counters = counters;
    const int chunks = detr_colsum_chunks(M, N);
    p.rows_per_cta = (M + chunks - 1) / chunks;
    dim3 grid((N + kColsPerCta - 1) / kColsPerCta, chunks);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0 && g_dtype == 0) epilogue_bwd_kernel<0, float><<<grid, kCsThreads, 0, st>>>(p);
    else if (mode == 0) epilogue_bwd_kernel<0, __nv_bfloat16><<<grid, kCsThreads, 0, st>>>(p);
    else if (g_dtype == 0) epilogue_bwd_kernel<1, float><<<grid, kCsThreads, 0, st>>>(p);
    else epilogue_bwd_kernel<1, __nv_bfloat16><<<grid, kCsThreads, 0, st>>>(p);
    DETR_CHECK_LAUNCH("epilogue_bwd");
    return 0;
}

// =========================================================================================================
// Backward entry of the prediction heads (detr/model.py:92-93: class_embedding, bbox_embedding(.).sigmoid()).
// The criterion hands back fp32 d_logits [rows][K] and d_boxes [rows][4]; the heads' input- and weight-gradient GEMMs
// (csrc/gemm.cu) want bf16 operands whose widths are whole MMA tiles.  One pass writes both:
//   dl16[rows][ld_l]  = bf16(d_logits), zero beyond column K
//   dz16[rows][ld_z]  = bf16(d_boxes * b * (1 - b)), b = the sigmoid output, zero beyond column 4
// One warp per row.
// =========================================================================================================
namespace detr {

__global__ void __launch_bounds__(256) heads_grad_prep_kernel(const float* __restrict__ d_logits, int K, const float* __restrict__ d_boxes,
                                                              const float* __restrict__ boxes, __nv_bfloat16* __restrict__ dl16, int ld_l,
                                                              __nv_bfloat16* __restrict__ dz16, int ld_z, int rows) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* g = d_logits + (int64_t)row * K;
    for (int c = lane * 2; c < ld_l; c += 64) {
        const float a = c < K ? __ldg(g + c) : 0.f, b = c + 1 < K ? __ldg(g + c + 1) : 0.f;
        *reinterpret_cast<__nv_bfloat162*>(dl16 + (int64_t)row * ld_l + c) = __floats2bfloat162_rn(a, b);
    }
    for (int c = lane * 2; c < ld_z; c += 64) {
        float v[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (c + i < 4) {
                const float s = __ldg(boxes + (int64_t)row * 4 + c + i);
                v[i] = __ldg(d_boxes + (int64_t)row * 4 + c + i) * s * (1.f - s);
            }
        *reinterpret_cast<__nv_bfloat162*>(dz16 + (int64_t)row * ld_z + c) = __floats2bfloat162_rn(v[0], v[1]);
    }
}

}  // namespace detr

extern "C" int detr_heads_grad_prep(const float* d_logits, int K, const float* d_boxes, const float* boxes, void* dl16, int ld_l,
                                    void* dz16, int ld_z, int rows, void* stream) {
    DETR_CHECK_ARG(d_logits && d_boxes && boxes && dl16 && dz16 && rows >= 1 && K >= 1, "heads_grad_prep: null argument");
    DETR_CHECK_ARG(ld_l >= K && ld_l % 2 == 0 && ld_z >= 4 && ld_z % 2 == 0 && ((uintptr_t)dl16 % 4) == 0 && ((uintptr_t)dz16 % 4) == 0,
                   "heads_grad_prep: ld_l >= K, ld_z >= 4, both even");
    detr::heads_grad_prep_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(d_logits, K, d_boxes, boxes,
                                                                                   reinterpret_cast<__nv_bfloat16*>(dl16), ld_l,
                                                                                   reinterpret_cast<__nv_bfloat16*>(dz16), ld_z, rows);
    DETR_CHECK_LAUNCH("heads_grad_prep");
    return 0;
}
