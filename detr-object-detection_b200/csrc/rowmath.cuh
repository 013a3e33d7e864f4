// Element-wise math shared by the row-op kernels (rowops.cu) and the GEMM epilogues (gemm.cu): tanh-GELU and its
// derivative (detr/model.py:406 nn.GELU(approximate="tanh")) and the counter-based dropout mask of the block tails
// (detr/model.py:355,408,410), indexed by the 8-element chunk (row * N/8 + column/8) of the (M, N) activation.
#pragma once
#include "tc.cuh"

namespace detr {

__device__ __forceinline__ float gelu_tanh_fwd(float a, float& t_out) {
    const float u = 0.7978845608028654f * (a + 0.044715f * a * a * a);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    t_out = t;
    return 0.5f * a * (1.f + t);
}
__device__ __forceinline__ float gelu_tanh_grad(float a) {
    float t;
    gelu_tanh_fwd(a, t);
    const float du = 0.7978845608028654f * (1.f + 3.f * 0.044715f * a * a);
    return 0.5f * (1.f + t) + 0.5f * a * (1.f - t * t) * du;
}

// keep[e] for the 8 elements of chunk `idx8` (thr2 = threshold * 0x00010001, threshold in 1/32768: tc::dropout_threshold)
__device__ __forceinline__ void ew_keep8(uint32_t key, uint32_t idx8, uint32_t thr2, bool* keep) {
    uint32_t st = tc::dropout_group_state(key, idx8);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const uint32_t t = tc::dropout_pair(st, thr2);
        keep[2 * e] = (t >> 15) & 1u; keep[2 * e + 1] = (t >> 31) & 1u;
    }
}
__device__ __forceinline__ uint32_t ew_key(uint64_t seed, const uint64_t* seed_ptr) {
    const uint64_t s = seed + (seed_ptr ? *seed_ptr : 0ull);
    return tc::mix32((uint32_t)s ^ tc::mix32((uint32_t)(s >> 32) + 0x9E3779B9u));
}

}  // namespace detr
