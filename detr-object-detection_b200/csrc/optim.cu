// Caller-side glue of the benchmark harness (detr/train.py:265-267: clip_grad_norm_(1.0) + AdamW.step): with parameters,
// gradients and both moments living in FLAT fp32 buffers, the optimizer is two HBM-speed launches instead of ATen's ~28
// (multi-tensor norm, stack, norm, clamp, multi-tensor mul, fused AdamW over ~330 tensors: 0.8 ms per step, 3x the traffic
// bound):
//   sumsq   partial sums of g^2 (one per CTA), folded in a fixed order by the last CTA -> total ||g||^2 (deterministic)
//   adamw   p, m, v updated in place with the clip coefficient min(1, max_norm / (||g|| + 1e-6)) applied to g on the fly
// Semantics of torch.optim.AdamW (decoupled weight decay, bias correction with the step counter read from device memory,
// so the launch is CUDA-graph replayable) and of torch.nn.utils.clip_grad_norm_ (L2, error_if_nonfinite=False).
#include "common.cuh"

namespace detr {

constexpr int kOptThreads = 256;

__global__ void __launch_bounds__(kOptThreads) sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial,
                                                            float* __restrict__ out, unsigned* __restrict__ counter, float* __restrict__ step) {
    __shared__ float red[kOptThreads / 32];
    __shared__ bool is_last;
    float acc = 0.f;
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kOptThreads) {
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[(n4 << 2) + threadIdx.x]; acc += v * v; }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < kOptThreads / 32; ++w) s += red[w];
        partial[blockIdx.x] = s;
        __threadfence();
        is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last) {   // fixed-order fold of <= 1184 partials by one warp-sized loop per thread, then the CTA
        float s = 0.f;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += kOptThreads) s += __ldcg(&partial[k]);
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int w = 0; w < kOptThreads / 32; ++w) t += red[w];
            *out = t;
            *counter = 0;
            // the optimizer's step counter advances only when the update will be applied (adamw_clip_kernel skips a step whose
            // gradient norm is not finite: a faulted batch must not poison the weights and the moments)
            if (step != nullptr && isfinite(t)) *step += 1.f;
        }
    }
}

__global__ void __launch_bounds__(kOptThreads) adamw_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                                 float* __restrict__ v, int64_t n, const float* __restrict__ lr_ptr, float beta1, float beta2,
                                                                 float eps, float weight_decay, const float* __restrict__ step,
                                                                 const float* __restrict__ sumsq, float max_norm, float grad_div) {
    if (sumsq != nullptr && !isfinite(*sumsq)) return;   // NaN / inf gradients (e.g. a faulted batch whose losses were poisoned): skip the update
    const float t = *step, lr = *lr_ptr;   // both in device memory: a captured CUDA graph follows the LR schedule and the step count
    const float bc1 = 1.f - powf(beta1, t), bc2_sqrt = sqrtf(1.f - powf(beta2, t));
    float coef = grad_div;                                      // 1 / world size folded in (data-parallel mean)
    if (sumsq != nullptr && max_norm > 0.f) {
        const float norm = sqrtf(*sumsq) * grad_div;
        coef *= fminf(1.f, max_norm / (norm + 1e-6f));
    }
    const float step_size = lr / bc1, decay = 1.f - lr * weight_decay;
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kOptThreads) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float* P = &pp.x; float* M = &mm.x; float* V = &vv.x; const float* G = &gg.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float ge = G[e] * coef;
            M[e] = beta1 * M[e] + (1.f - beta1) * ge;
            V[e] = beta2 * V[e] + (1.f - beta2) * ge * ge;
            const float denom = sqrtf(V[e]) / bc2_sqrt + eps;
            P[e] = P[e] * decay - step_size * (M[e] / denom);
        }
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (n4 << 2) + threadIdx.x;
        const float ge = g[i] * coef;
        const float me = beta1 * m[i] + (1.f - beta1) * ge, ve = beta2 * v[i] + (1.f - beta2) * ge * ge;
        m[i] = me; v[i] = ve;
        p[i] = p[i] * decay - step_size * (me / (sqrtf(ve) / bc2_sqrt + eps));
    }
}

}  // namespace detr

using namespace detr;

extern "C" int detr_sumsq_grid(long long n) {
    long long g = (n / 4 + kOptThreads * 8 - 1) / (kOptThreads * 8);
    if (g > 148 * 8) g = 148 * 8;
    return (int)(g < 1 ? 1 : g);
}

extern "C" int detr_sumsq_f32(const float* g, long long n, float* partial, float* out, uint32_t* counter, float* step, void* stream) {
    DETR_CHECK_ARG(g != nullptr && n >= 1 && ((uintptr_t)g % 16) == 0 && partial && out && counter, "sumsq: bad arguments");
    sumsq_kernel<<<detr_sumsq_grid(n), kOptThreads, 0, (cudaStream_t)stream>>>(g, n, partial, out, counter, step);
    DETR_CHECK_LAUNCH("sumsq");
    return 0;
}

extern "C" int detr_adamw_clip_f32(float* p, const float* g, float* m, float* v, long long n, const float* lr, float beta1, float beta2, float eps,
                                   float weight_decay, const float* step, const float* sumsq, float max_norm, float grad_div, void* stream) {
    DETR_CHECK_ARG(p && g && m && v && step && lr && n >= 1, "adamw_clip: null pointer or empty range");
    DETR_CHECK_ARG(((uintptr_t)p % 16) == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)m % 16) == 0 && ((uintptr_t)v % 16) == 0, "adamw_clip: 16-byte alignment");
    long long grid = (n / 4 + kOptThreads * 4 - 1) / (kOptThreads * 4);
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    adamw_clip_kernel<<<(int)grid, kOptThreads, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, sumsq, max_norm, grad_div);
    DETR_CHECK_LAUNCH("adamw_clip");
    return 0;
}
