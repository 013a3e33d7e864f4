// Fused SetCriterion: forward (2 launches; 3 on the fallback path) and backward (1 launch) for ALL decoder layers at once.
//
// Replaces detr/loss.py:198-231: per layer { loss_labels (57-95), loss_cardinality (97-121), loss_boxes (123-164) },
// i.e. ~60 ATen kernels + CPU-index -> CUDA-index copies per layer, by
//   criterion_fwd_dense_kernel one CTA per (image, layer), dense logits blocks with K % 4 == 0 (the DETR case): the 36.8 KB
//                              block is staged by cp.async.bulk + mbarrier; under the copy the assignment (idx_q, idx_gt) is
//                              expanded to dense per-query arrays -- target class (K-1 = "no object") and matched target box
//                              (NaN = unmatched) -- and the L1 / GIoU terms are computed; then ONE THREAD PER QUERY ROW walks
//                              its row (LDS.128, conflict-free): log-sum-exp, weighted NLL numerator/denominator, arg-max
//                              (cardinality, class_error).  Per-problem partial sums, no float atomics.
//   criterion_expand_kernel +  fallback for strided rows or K % 4 != 0: assignment expanded by a small kernel, then one warp
//   criterion_fwd_kernel<C>    per row with the logits of 13 rows in registers (C = 32-wide column chunks).
//   criterion_finalize_kernel  fixed-order reduction over images -> the L x 5 loss table (deterministic).
//   criterion_bwd_dense_kernel same staging; each thread turns its row into c * (softmax - onehot) IN PLACE and the block
//                              leaves by cp.async.bulk stores; grad_boxes analytic (L1 + GIoU) from the dense per-query
//                              arrays: no indices, no offsets, no dependent loads.  Fallback: criterion_bwd_kernel<vec>.
#include <math_constants.h>

#include "common.cuh"
#include "tc.cuh"

namespace detr {

constexpr int kCritThreads = 256;
constexpr int kCritWarps = kCritThreads / 32;
constexpr int kPartials = 8;  // {sum w*nll, sum w, n_pred_nonempty, n_correct, l1_sum, giou_sum, n_pairs, unused}
constexpr int kFwdRows = 13;  // query rows a warp keeps in registers at once: 8 warps x 13 rows cover Q = 100 in one batch

struct CritParams {
    const float* logits; int64_t lg_sb, lg_sl, lg_sq;
    const float* boxes;  int64_t bx_sb, bx_sl, bx_sq;
    const int64_t* gt_labels; const float* gt_boxes; const int32_t* gt_off; const int32_t* match_off;
    const int64_t* idx_q; const int64_t* idx_gt;
    const float* class_weight; const float* num_boxes;
    int B, L, Q, K;
    float w_ce, w_l1, w_giou;
    float* partials; float* lse; int32_t* tgt; float* tbox; float* wsum; float* losses;
    int32_t* status;
    // backward only
    const float* grad_losses; float* grad_logits; float* grad_boxes;
};

struct Box4 { float x1, y1, x2, y2; };

__device__ __forceinline__ Box4 cxcywh_to_xyxy(float4 c) {
    Box4 r;
    r.x1 = __fsub_rn(c.x, __fdiv_rn(c.z, 2.f));
    r.y1 = __fsub_rn(c.y, __fdiv_rn(c.w, 2.f));
    r.x2 = __fadd_rn(c.z, r.x1);
    r.y2 = __fadd_rn(c.w, r.y1);
    return r;
}

constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float redux_max_f32(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));   // CREDUX.MAX.F32: one instruction per warp
    return r;
}

// Pair losses of one query against its matched target box (XYXY): L1 against cxcywh(target) (detr/loss.py:149-156) and the
// GIoU loss with eps (torchvision giou_loss.py:47-62, _utils.py:87-106).
__device__ __forceinline__ void pair_losses(float4 s, float4 t, float& l1, float& gi) {
    const float tw = __fsub_rn(t.z, t.x), th = __fsub_rn(t.w, t.y);
    const float tcx = __fdiv_rn(__fadd_rn(__fmul_rn(t.x, 2.f), tw), 2.f);
    const float tcy = __fdiv_rn(__fadd_rn(__fmul_rn(t.y, 2.f), th), 2.f);
    l1 = fabsf(s.x - tcx) + fabsf(s.y - tcy) + fabsf(s.z - tw) + fabsf(s.w - th);
    const Box4 a = cxcywh_to_xyxy(s);
    const float eps = 1e-7f;
    const float ix1 = fmaxf(a.x1, t.x), iy1 = fmaxf(a.y1, t.y), ix2 = fminf(a.x2, t.z), iy2 = fminf(a.y2, t.w);
    const float inter = (iy2 > iy1 && ix2 > ix1) ? __fmul_rn(ix2 - ix1, iy2 - iy1) : 0.f;
    const float uni = __fsub_rn(__fadd_rn(__fmul_rn(a.x2 - a.x1, a.y2 - a.y1), __fmul_rn(t.z - t.x, t.w - t.y)), inter);
    const float iou = __fdiv_rn(inter, uni + eps);
    const float hull = __fmul_rn(fmaxf(a.x2, t.z) - fminf(a.x1, t.x), fmaxf(a.y2, t.w) - fminf(a.y1, t.y));
    gi = 1.f - (iou - __fdiv_rn(hull - uni, hull + eps));
}

constexpr int kExpandThreads = 128;

// (idx_q, idx_gt) of a problem -> dense per-query target class and matched target box (detr/loss.py:79-85, 144-147).
__global__ void __launch_bounds__(kExpandThreads) criterion_expand_kernel(const CritParams p) {
    extern __shared__ int s_g[];   // [Q] matched gt index (global row of the packed targets) or -1
    const int tid = threadIdx.x;
    const int b = blockIdx.x / p.L, l = blockIdx.x % p.L;
    const int Q = p.Q, K = p.K;
    const int g0 = p.gt_off[b], M = p.gt_off[b + 1] - g0;
    const int n = min(Q, M);
    const int64_t moff = (int64_t)p.L * p.match_off[b] + (int64_t)l * n;
    for (int q = tid; q < Q; q += kExpandThreads) s_g[q] = -1;
    __syncthreads();
    for (int k = tid; k < n; k += kExpandThreads) {
        const int64_t q = p.idx_q[moff + k], g = p.idx_gt[moff + k];
        if (q < 0 || q >= Q || g < 0 || g >= M) continue;  // poisoned by a failed assignment: status already set
        s_g[q] = g0 + (int)g;
    }
    __syncthreads();
    int32_t* tgt = p.tgt + (int64_t)blockIdx.x * Q;
    float4* tbox = reinterpret_cast<float4*>(p.tbox) + (int64_t)blockIdx.x * Q;
    for (int q = tid; q < Q; q += kExpandThreads) {
        const int g = s_g[q];
        int t = K - 1;
        float4 tb = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
        if (g >= 0) {
            int64_t lab = p.gt_labels[g];
            tb = *reinterpret_cast<const float4*>(p.gt_boxes + (int64_t)g * 4);
            if (lab < 0 || lab >= K) { atomicOr(p.status, DETR_ST_BAD_LABEL); lab = K - 1; }
            t = (int)lab;
            if (!(tb.x == tb.x)) tb.x = 0.f;   // NaN x1 is the "unmatched" flag; a NaN in the data has already raised
                                               // DETR_ST_DEGENERATE_BOX in the matcher, which poisons the losses
        }
        tgt[q] = t;
        tbox[q] = tb;
    }
}

// One CTA per (image, layer).  kC = 32-wide column chunks of a logits row held in registers (K <= 32 * kC); kC == 0 is the
// generic path for wider rows (one row per warp at a time, logits re-read from L1/L2).
//
// A problem is 36.8 KB of logits (Q=100, K=92): all of its loads (13 rows x kC registers per lane, then class / boxes of
// one query per thread) are in flight before the first use, and nothing in the kernel depends on a loaded index.
template <int kC>
__global__ void __launch_bounds__(kCritThreads, 3) criterion_fwd_kernel(const CritParams p) {
    extern __shared__ int s_dyn[];  // [Q] row max | [Q] row arg-max | [Q] row sum of exp
    __shared__ float red[kCritWarps][kPartials];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x / p.L, l = blockIdx.x % p.L;
    const int Q = p.Q, K = p.K;
    float* s_mx = reinterpret_cast<float*>(s_dyn);
    int* s_am = s_dyn + Q;
    float* s_sum = reinterpret_cast<float*>(s_dyn + 2 * Q);
    const float* lg = p.logits + b * p.lg_sb + l * p.lg_sl;
    const float* bx = p.boxes + b * p.bx_sb + l * p.bx_sl;
    const int32_t* tgt = p.tgt + (int64_t)blockIdx.x * Q;
    const float4* tbox = reinterpret_cast<const float4*>(p.tbox) + (int64_t)blockIdx.x * Q;
    constexpr int kCr = kC > 0 ? kC : 1;

    float l1 = 0.f, gi = 0.f, npairs = 0.f;
    for (int base = 0; base < Q; base += kCritWarps * kFwdRows) {
        float v[kFwdRows][kCr];
        if (kC > 0) {
#pragma unroll
            for (int r = 0; r < kFwdRows; ++r) {
                const int q = base + r * kCritWarps + warp;
                const float* row = lg + (int64_t)q * p.lg_sq;
#pragma unroll
                for (int c = 0; c < kCr; ++c) {
                    const int k = lane + 32 * c;
                    v[r][c] = (q < Q && k < K) ? __ldg(row + k) : -CUDART_INF_F;
                }
            }
        }
        if (base == 0) {
            // box losses of the matched queries, one query per thread (detr/loss.py:144-162)
            for (int q = tid; q < Q; q += kCritThreads) {
                const float4 t = tbox[q];
                const float4 s = *reinterpret_cast<const float4*>(bx + (int64_t)q * p.bx_sq);
                if (t.x == t.x) {
                    float a, g;
                    pair_losses(s, t, a, g);
                    l1 += a; gi += g; npairs += 1.f;
                }
            }
        }
        // row max, arg-max (lowest index wins ties: torch.argmax / topk on distinct values is unaffected) and sum of
        // exp(x - max).  Straight-line over the 13 rows (out-of-range rows and columns hold -inf and only cost issue slots):
        // the compiler interleaves the rows' CREDUX / SHFL latencies.  exp(x - m) = ex2(x * log2e - m * log2e): one FFMA and
        // one MUFU per logit (the kernel was issue-bound on libdevice's 9-instruction expf: ncu, profiles/r01_criterion_ncu.md).
        if (kC > 0) {
#pragma unroll
            for (int r = 0; r < kFwdRows; ++r) {
                const int q = base + r * kCritWarps + warp;
                float mx = v[r][0];
#pragma unroll
                for (int c = 1; c < kCr; ++c) mx = fmaxf(mx, v[r][c]);
                mx = redux_max_f32(mx);
                const float nb = -mx * kLog2e;
                int am = 0x7fffffff;
                float sum = 0.f;
#pragma unroll
                for (int c = 0; c < kCr; ++c) {
                    am = v[r][c] == mx ? min(am, lane + 32 * c) : am;
                    sum += ex2_approx(fmaf(v[r][c], kLog2e, nb));
                }
                am = __reduce_min_sync(FULL_MASK, am);
                sum = warp_sum(sum);
                if (lane == 0 && q < Q) { s_mx[q] = mx; s_sum[q] = sum; s_am[q] = am; }
            }
        } else {
            for (int r = 0; r < kFwdRows; ++r) {
                const int q = base + r * kCritWarps + warp;
                if (q >= Q) break;   // warp-uniform
                const float* row = lg + (int64_t)q * p.lg_sq;
                float mx = -CUDART_INF_F, sum = 0.f;
                int am = 0x7fffffff;
                for (int k = lane; k < K; k += 32) mx = fmaxf(mx, row[k]);
                mx = redux_max_f32(mx);
                const float nb = -mx * kLog2e;
                for (int k = lane; k < K; k += 32) {
                    const float x = row[k];
                    am = x == mx ? min(am, k) : am;
                    sum += ex2_approx(fmaf(x, kLog2e, nb));
                }
                am = __reduce_min_sync(FULL_MASK, am);
                sum = warp_sum(sum);
                if (lane == 0) { s_mx[q] = mx; s_sum[q] = sum; s_am[q] = am; }
            }
        }
    }
    __syncthreads();

    // one thread per query row: weighted NLL of its target class, cardinality and class_error counts
    float wnll = 0.f, wsum = 0.f, nonempty = 0.f, correct = 0.f;
    float* lse_out = p.lse + (int64_t)blockIdx.x * Q;
    for (int q = tid; q < Q; q += kCritThreads) {
        const int t = tgt[q], am = s_am[q];
        const float lse = fmaf(lg2_approx(s_sum[q]), kLn2, s_mx[q]);
        const float w = p.class_weight[t];
        const float xt = lg[(int64_t)q * p.lg_sq + t];
        wnll += w * (lse - xt);
        wsum += w;
        nonempty += (am != K - 1) ? 1.f : 0.f;
        const float tx = tbox[q].x;
        if (tx == tx) correct += (am == t) ? 1.f : 0.f;   // matched queries only (detr/loss.py:93)
        lse_out[q] = lse;
    }
    const float vals[7] = {wnll, wsum, nonempty, correct, l1, gi, npairs};
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const float s = warp_sum(vals[k]);
        if (lane == 0) red[warp][k] = s;
    }
    __syncthreads();
    if (tid < kPartials) {
        float s = 0.f;
        if (tid < 7) {
#pragma unroll
            for (int w = 0; w < kCritWarps; ++w) s += red[w][tid];   // fixed order: deterministic
        }
        p.partials[(int64_t)blockIdx.x * kPartials + tid] = s;
    }
}

// =========================================================================================================
// Dense path: the (Q, K) logits block of a problem is contiguous and 16-byte aligned.  It is staged in shared memory by
// bulk asynchronous copies (cp.async.bulk, the 1-D TMA path) issued by one thread and tracked by an mbarrier, so the
// bytes in flight cost no registers: 5 problems (5 x 37 KB) are resident per SM instead of 3 with register staging, and
// the small per-query loads and the box math run underneath the copy.
// =========================================================================================================
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
// whole block in <= 4 copies (each a multiple of 16 bytes) so that several requests are in flight per CTA
__device__ __forceinline__ void stage_block(float* s_dst, const float* g_src, uint32_t bytes, uint64_t* bar) {
    tc::mbar_expect_tx(bar, bytes);
    const uint32_t chunk = ((bytes / 4 + 15u) / 16u) * 16u;
    for (uint32_t off = 0; off < bytes; off += chunk)
        bulk_load_1d(reinterpret_cast<char*>(s_dst) + off, reinterpret_cast<const char*>(g_src) + off, min(chunk, bytes - off), bar);
}

constexpr int kRowThreads = 128;   // one query row per thread: Q = 100 fills 4 warps, one per scheduler
constexpr int kRowWarps = kRowThreads / 32;

// One CTA per (image, layer), one THREAD per query row.  The warp-per-row kernels above spend ~140 instructions per row on
// predicated loads and three warp reductions (ncu: 23 M warp instructions at config 3 = issue-bound at ~20 us); here a row
// is walked by one thread as K/4 LDS.128 (a row is K/4 float4 long: conflict-free when K/4 is odd, e.g. K = 92), no
// reduction at all, ~6x fewer instructions.  The assignment is expanded by the same CTA while the copy is in flight
// (offsets -> indices -> labels / boxes is a chain of 4 dependent loads, about as long as the copy itself), so the forward
// call is 2 launches instead of 3.
__global__ void __launch_bounds__(kRowThreads) criterion_fwd_dense_kernel(const CritParams p) {
    extern __shared__ __align__(16) int s_dyn[];  // [Q*K] logits | [Q] matched gt row of the packed targets or -1
    __shared__ float red[kRowWarps][kPartials];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x / p.L, l = blockIdx.x % p.L;
    const int Q = p.Q, K = p.K;
    float* s_lg = reinterpret_cast<float*>(s_dyn);
    int* s_g = s_dyn + Q * K;
    const float* lg = p.logits + b * p.lg_sb + l * p.lg_sl;
    const float* bx = p.boxes + b * p.bx_sb + l * p.bx_sl;
    int32_t* tgt = p.tgt + (int64_t)blockIdx.x * Q;
    float4* tbox = reinterpret_cast<float4*>(p.tbox) + (int64_t)blockIdx.x * Q;
    float* lse_out = p.lse + (int64_t)blockIdx.x * Q;

    pdl_trigger();   // the finalize launch may be scheduled early; it waits for this grid before reading the partial sums
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    for (int q = tid; q < Q; q += kRowThreads) s_g[q] = -1;
    __syncthreads();
    if (tid == 0) stage_block(s_lg, lg, (uint32_t)(Q * K) * 4u, &bar);

    // under the copy: (idx_q, idx_gt) -> matched gt row per query (detr/loss.py:79-85, 144-147)
    const int g0 = p.gt_off[b], M = p.gt_off[b + 1] - g0;
    const int n = min(Q, M);
    const int64_t moff = (int64_t)p.L * p.match_off[b] + (int64_t)l * n;
    for (int k = tid; k < n; k += kRowThreads) {
        const int64_t q = p.idx_q[moff + k], g = p.idx_gt[moff + k];
        if (q < 0 || q >= Q || g < 0 || g >= M) continue;  // poisoned by a failed assignment: status already set
        s_g[q] = g0 + (int)g;
    }
    __syncthreads();

    float wnll = 0.f, wsum = 0.f, nonempty = 0.f, correct = 0.f, l1 = 0.f, gi = 0.f, npairs = 0.f;
    bool landed = false;
    const int K4 = K >> 2;
    for (int q = tid; q < Q; q += kRowThreads) {
        // still under the copy: target class / box of this query, class weight, box losses (detr/loss.py:144-162)
        const int g = s_g[q];
        const float4 s = *reinterpret_cast<const float4*>(bx + (int64_t)q * p.bx_sq);
        int t = K - 1;
        float4 tb = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
        if (g >= 0) {
            int64_t lab = p.gt_labels[g];
            tb = *reinterpret_cast<const float4*>(p.gt_boxes + (int64_t)g * 4);
            if (lab < 0 || lab >= K) { atomicOr(p.status, DETR_ST_BAD_LABEL); lab = K - 1; }
            t = (int)lab;
            if (!(tb.x == tb.x)) tb.x = 0.f;   // NaN x1 is the "unmatched" flag; a NaN in the data has already raised
                                               // DETR_ST_DEGENERATE_BOX in the matcher, which poisons the losses
        }
        const float w = p.class_weight[t];
        tgt[q] = t;
        tbox[q] = tb;
        if (g >= 0) {
            float a, gl;
            pair_losses(s, tb, a, gl);
            l1 += a; gi += gl; npairs += 1.f;
        }
        if (!landed) { tc::mbar_wait(&bar, 0); landed = true; }

        // row max; then arg-max (lowest index wins ties: walked from the last column down) and sum of
        // exp(x - max) = ex2(x * log2e - max * log2e)
        const float4* row = reinterpret_cast<const float4*>(s_lg + q * K);
        float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F;
#pragma unroll 4
        for (int j = 0; j < K4; ++j) {
            const float4 x = row[j];
            m0 = fmaxf(m0, x.x); m1 = fmaxf(m1, x.y); m2 = fmaxf(m2, x.z); m3 = fmaxf(m3, x.w);
        }
        const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        const float nb = -mx * kLog2e;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int am = 0;
#pragma unroll 4
        for (int j = K4 - 1; j >= 0; --j) {
            const float4 x = row[j];
            am = x.w == mx ? 4 * j + 3 : am;
            am = x.z == mx ? 4 * j + 2 : am;
            am = x.y == mx ? 4 * j + 1 : am;
            am = x.x == mx ? 4 * j : am;
            s0 += ex2_approx(fmaf(x.x, kLog2e, nb)); s1 += ex2_approx(fmaf(x.y, kLog2e, nb));
            s2 += ex2_approx(fmaf(x.z, kLog2e, nb)); s3 += ex2_approx(fmaf(x.w, kLog2e, nb));
        }
        const float lse = fmaf(lg2_approx((s0 + s1) + (s2 + s3)), kLn2, mx);
        wnll += w * (lse - s_lg[q * K + t]);
        wsum += w;
        nonempty += (am != K - 1) ? 1.f : 0.f;
        if (g >= 0) correct += (am == t) ? 1.f : 0.f;   // matched queries only (detr/loss.py:93)
        lse_out[q] = lse;
    }
    if (!landed) tc::mbar_wait(&bar, 0);   // the copy must have landed before the CTA may retire

    const float vals[7] = {wnll, wsum, nonempty, correct, l1, gi, npairs};
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const float v = warp_sum(vals[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (tid < kPartials) {
        float v = 0.f;
        if (tid < 7) {
#pragma unroll
            for (int w = 0; w < kRowWarps; ++w) v += red[w][tid];   // fixed order: deterministic
        }
        p.partials[(int64_t)blockIdx.x * kPartials + tid] = v;
    }
}

// one CTA per layer; images are folded in a fixed order (thread -> warp -> CTA) -> bitwise reproducible losses
constexpr int kFinThreads = 256;
__global__ void __launch_bounds__(kFinThreads) criterion_finalize_kernel(const CritParams p) {
    __shared__ float red[kFinThreads / 32][8];
    const int l = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_wait();
    float acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int b = tid; b < p.B; b += kFinThreads) {   // one image per thread: all loads of the launch are in flight at once
        const float* s = p.partials + ((int64_t)b * p.L + l) * kPartials;
        const float m = (float)(p.gt_off[b + 1] - p.gt_off[b]);
        acc[0] += s[0]; acc[1] += s[1]; acc[2] += fabsf(s[2] - m); acc[3] += s[3]; acc[4] += s[4]; acc[5] += s[5]; acc[6] += s[6];
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const float v = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < kFinThreads / 32; ++w) v += red[w][k];
            acc[k] = v;
        }
        const float nb = p.num_boxes ? *p.num_boxes : fmaxf((float)p.gt_off[p.B], 1.f);  // detr/loss.py:142
        float* o = p.losses + l * 5;
        o[0] = p.w_ce * (acc[0] / acc[1]);                 // weighted mean CE (detr/loss.py:90-91)
        o[1] = acc[2] / (float)p.B;                        // cardinality error (detr/loss.py:121)
        o[2] = p.w_l1 * acc[4] / nb;                       // detr/loss.py:152-156
        o[3] = p.w_giou * acc[5] / nb;                     // detr/loss.py:158-162
        o[4] = acc[6] > 0.f ? 100.f - acc[3] * (100.f / acc[6]) : 100.f;  // detr/loss.py:93, detr/utils.py:100-116
        p.wsum[l] = acc[1];
        // A data fault (degenerate box, NaN cost, infeasible assignment) raises in the reference; here it
        // poisons the losses so that the step cannot silently train on a wrong assignment.
        if (*p.status != 0) { o[0] = o[1] = o[2] = o[3] = o[4] = CUDART_NAN_F; }
    }
}

__device__ __forceinline__ float step_gt(float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); }

constexpr int kBwdVec = 9;  // float4 per thread in flight on the dense path: 256 x 9 x 4 covers Q*K = 9 200 in one batch

// Gradient of (s_l1 * L1 + s_gi * GIoU loss) of one matched pair w.r.t. the predicted cxcywh box.
__device__ __forceinline__ float4 pair_grads(float4 s, float4 t, float s_l1, float s_gi) {
    const float tw = __fsub_rn(t.z, t.x), th = __fsub_rn(t.w, t.y);
    const float tcx = __fdiv_rn(__fadd_rn(__fmul_rn(t.x, 2.f), tw), 2.f);
    const float tcy = __fdiv_rn(__fadd_rn(__fmul_rn(t.y, 2.f), th), 2.f);
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
    float dcx = s_l1 * sgn(s.x - tcx), dcy = s_l1 * sgn(s.y - tcy), dw = s_l1 * sgn(s.z - tw), dh = s_l1 * sgn(s.w - th);

    const Box4 a = cxcywh_to_xyxy(s);
    const float eps = 1e-7f;
    const float ix1 = fmaxf(a.x1, t.x), iy1 = fmaxf(a.y1, t.y), ix2 = fminf(a.x2, t.z), iy2 = fminf(a.y2, t.w);
    const bool ok = (iy2 > iy1) && (ix2 > ix1);
    const float iw = ix2 - ix1, ih = iy2 - iy1;
    const float inter = ok ? iw * ih : 0.f;
    const float aw = a.x2 - a.x1, ah = a.y2 - a.y1;
    const float uni = aw * ah + (t.z - t.x) * (t.w - t.y) - inter;
    const float hx1 = fminf(a.x1, t.x), hy1 = fminf(a.y1, t.y), hx2 = fmaxf(a.x2, t.z), hy2 = fmaxf(a.y2, t.w);
    const float hw = hx2 - hx1, hh = hy2 - hy1;
    const float hull = hw * hh;
    // partial derivatives w.r.t. (x1, y1, x2, y2) of the predicted box; max/min split ties 0.5/0.5 like autograd
    float dI[4] = {0.f, 0.f, 0.f, 0.f};
    if (ok) {
        dI[0] = -ih * step_gt(a.x1, t.x); dI[1] = -iw * step_gt(a.y1, t.y);
        dI[2] = ih * step_gt(t.z, a.x2);  dI[3] = iw * step_gt(t.w, a.y2);
    }
    const float dA[4] = {-ah, -aw, ah, aw};
    const float dH[4] = {-hh * step_gt(t.x, a.x1), -hw * step_gt(t.y, a.y1), hh * step_gt(a.x2, t.z), hw * step_gt(a.y2, t.w)};
    const float ue = uni + eps, he = hull + eps;
    float dxy[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float dU = dA[c] - dI[c];
        const float d_iou = dI[c] / ue - inter * dU / (ue * ue);
        const float d_pen = (dH[c] - dU) / he - (hull - uni) * dH[c] / (he * he);
        dxy[c] = s_gi * (-d_iou + d_pen);
    }
    // chain through x1 = cx - w/2, x2 = w + x1  =>  d/dcx = dx1 + dx2 ; d/dw = -dx1/2 + dx2/2
    dcx += dxy[0] + dxy[2]; dcy += dxy[1] + dxy[3];
    dw += 0.5f * (dxy[2] - dxy[0]); dh += 0.5f * (dxy[3] - dxy[1]);
    return make_float4(dcx, dcy, dw, dh);
}

// kVec: the (Q, K) logits block of a problem is dense, 16-byte aligned and K % 4 == 0 -> walked as flat float4 with every
// load of the block in flight before the first use; otherwise 4 rows per warp at a time.
template <bool kVec>
__global__ void __launch_bounds__(kCritThreads, 3) criterion_bwd_kernel(const CritParams p) {
    extern __shared__ int s_dyn[];  // kVec: [Q] coefficient | [Q] row log-sum-exp | [Q] target class
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x / p.L, l = blockIdx.x % p.L;
    const int Q = p.Q, K = p.K;
    const float* lg = p.logits + b * p.lg_sb + l * p.lg_sl;
    const float* bx = p.boxes + b * p.bx_sb + l * p.bx_sl;
    float* dlg = p.grad_logits + (int64_t)blockIdx.x * Q * K;
    const int n4 = (Q * K) >> 2;
    float4 x[kBwdVec];
    if (kVec) {
#pragma unroll
        for (int i = 0; i < kBwdVec; ++i) {
            const int e = tid + i * kCritThreads;
            if (e < n4) x[i] = __ldg(reinterpret_cast<const float4*>(lg) + e);
        }
    }
    const float g_ce = p.grad_losses[l * 5 + 0], g_l1 = p.grad_losses[l * 5 + 2], g_gi = p.grad_losses[l * 5 + 3];
    const float nb = p.num_boxes ? *p.num_boxes : fmaxf((float)p.gt_off[p.B], 1.f);
    const float ce_scale = g_ce * p.w_ce / p.wsum[l];   // d CE / d logits = g * w_ce * w[t]/W * (softmax - onehot)
    const float s_l1 = g_l1 * p.w_l1 / nb, s_gi = g_gi * p.w_giou / nb;
    const float* lse = p.lse + (int64_t)blockIdx.x * Q;
    const int32_t* tgt = p.tgt + (int64_t)blockIdx.x * Q;
    const float4* tbox = reinterpret_cast<const float4*>(p.tbox) + (int64_t)blockIdx.x * Q;
    float4* dbx = reinterpret_cast<float4*>(p.grad_boxes + (int64_t)blockIdx.x * Q * 4);

    // ---- per-query coefficients (dense path) and d boxes: zero for unmatched queries, analytic L1 + GIoU otherwise ----
    for (int q = tid; q < Q; q += kCritThreads) {
        const float4 t = tbox[q];
        const float4 s = *reinterpret_cast<const float4*>(bx + (int64_t)q * p.bx_sq);
        if (kVec) {
            float* s_cc = reinterpret_cast<float*>(s_dyn);
            const int tc = tgt[q];
            s_dyn[2 * Q + q] = tc;
            s_cc[Q + q] = lse[q];
            s_cc[q] = ce_scale * p.class_weight[tc];
        }
        dbx[q] = (t.x == t.x) ? pair_grads(s, t, s_l1, s_gi) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (kVec) {
        __syncthreads();
        const float* s_cc = reinterpret_cast<const float*>(s_dyn);
        const float* s_ls = s_cc + Q;
        const int* s_tt = s_dyn + 2 * Q;
        // element 4e of the block is (row q, column k); consecutive float4 of a thread are 4 * 256 elements apart: one
        // division per thread, then (q, k) advance incrementally
        const int step_q = (4 * kCritThreads) / K, step_k = (4 * kCritThreads) - step_q * K;
        int q = (4 * tid) / K, k = 4 * tid - q * K;
        for (int base = 0; base < n4; base += kBwdVec * kCritThreads) {
            if (base > 0) {
#pragma unroll
                for (int i = 0; i < kBwdVec; ++i) {
                    const int e = base + tid + i * kCritThreads;
                    if (e < n4) x[i] = __ldg(reinterpret_cast<const float4*>(lg) + e);
                }
            }
#pragma unroll
            for (int i = 0; i < kBwdVec; ++i) {
                const int e = base + tid + i * kCritThreads;
                if (e < n4) {   // K % 4 == 0: the four elements share a row
                    const float c = s_cc[q], nbias = -s_ls[q] * kLog2e;
                    const int t = s_tt[q] - k;
                    float4 o;
                    o.x = c * (ex2_approx(fmaf(x[i].x, kLog2e, nbias)) - (t == 0 ? 1.f : 0.f));
                    o.y = c * (ex2_approx(fmaf(x[i].y, kLog2e, nbias)) - (t == 1 ? 1.f : 0.f));
                    o.z = c * (ex2_approx(fmaf(x[i].z, kLog2e, nbias)) - (t == 2 ? 1.f : 0.f));
                    o.w = c * (ex2_approx(fmaf(x[i].w, kLog2e, nbias)) - (t == 3 ? 1.f : 0.f));
                    reinterpret_cast<float4*>(dlg)[e] = o;
                }
                q += step_q; k += step_k;
                if (k >= K) { k -= K; ++q; }
            }
        }
    } else {
        constexpr int kRows = 4, kMaxK = 4;
        const bool reg_path = K <= 32 * kMaxK;
        for (int q0 = warp; q0 < Q; q0 += kCritWarps * kRows) {   // loads of 4 rows in flight per warp
            float v[kRows][kMaxK], ls[kRows], cc[kRows];
            int tt[kRows];
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const int q = q0 + r * kCritWarps;
                const float* row = lg + (int64_t)q * p.lg_sq;
                tt[r] = q < Q ? tgt[q] : 0;
                ls[r] = q < Q ? lse[q] : 0.f;
#pragma unroll
                for (int c = 0; c < kMaxK; ++c) {
                    const int k = lane + 32 * c;
                    v[r][c] = (reg_path && q < Q && k < K) ? row[k] : 0.f;
                }
            }
#pragma unroll
            for (int r = 0; r < kRows; ++r) cc[r] = ce_scale * p.class_weight[tt[r]];
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const int q = q0 + r * kCritWarps;
                if (q >= Q) break;
                if (reg_path) {
#pragma unroll
                    for (int c = 0; c < kMaxK; ++c) {
                        const int k = lane + 32 * c;
                        if (k < K) dlg[(int64_t)q * K + k] = cc[r] * (expf(v[r][c] - ls[r]) - (k == tt[r] ? 1.f : 0.f));
                    }
                } else {
                    const float* row = lg + (int64_t)q * p.lg_sq;
                    for (int k = lane; k < K; k += 32) dlg[(int64_t)q * K + k] = cc[r] * (expf(row[k] - ls[r]) - (k == tt[r] ? 1.f : 0.f));
                }
            }
        }
    }
}

// pair_grads with the 16 IEEE divisions folded into two reciprocals (the divisions were a third of the backward kernel's
// instructions: ncu, profiles/r01_criterion_ncu_v2.md); same derivative, differences at the 1e-7 relative level.
__device__ __forceinline__ float4 pair_grads_fast(float4 s, float4 t, float s_l1, float s_gi) {
    const float tw = __fsub_rn(t.z, t.x), th = __fsub_rn(t.w, t.y);
    const float tcx = __fmul_rn(__fadd_rn(__fmul_rn(t.x, 2.f), tw), 0.5f);
    const float tcy = __fmul_rn(__fadd_rn(__fmul_rn(t.y, 2.f), th), 0.5f);
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
    float dcx = s_l1 * sgn(s.x - tcx), dcy = s_l1 * sgn(s.y - tcy), dw = s_l1 * sgn(s.z - tw), dh = s_l1 * sgn(s.w - th);
    Box4 a;
    a.x1 = __fsub_rn(s.x, __fmul_rn(s.z, 0.5f)); a.y1 = __fsub_rn(s.y, __fmul_rn(s.w, 0.5f));
    a.x2 = __fadd_rn(s.z, a.x1); a.y2 = __fadd_rn(s.w, a.y1);
    const float eps = 1e-7f;
    const float ix1 = fmaxf(a.x1, t.x), iy1 = fmaxf(a.y1, t.y), ix2 = fminf(a.x2, t.z), iy2 = fminf(a.y2, t.w);
    const bool ok = (iy2 > iy1) && (ix2 > ix1);
    const float iw = ix2 - ix1, ih = iy2 - iy1;
    const float inter = ok ? iw * ih : 0.f;
    const float aw = a.x2 - a.x1, ah = a.y2 - a.y1;
    const float uni = aw * ah + (t.z - t.x) * (t.w - t.y) - inter;
    const float hx1 = fminf(a.x1, t.x), hy1 = fminf(a.y1, t.y), hx2 = fmaxf(a.x2, t.z), hy2 = fmaxf(a.y2, t.w);
    const float hw = hx2 - hx1, hh = hy2 - hy1;
    const float hull = hw * hh;
    float dI[4] = {0.f, 0.f, 0.f, 0.f};
    if (ok) {
        dI[0] = -ih * step_gt(a.x1, t.x); dI[1] = -iw * step_gt(a.y1, t.y);
        dI[2] = ih * step_gt(t.z, a.x2);  dI[3] = iw * step_gt(t.w, a.y2);
    }
    const float dA[4] = {-ah, -aw, ah, aw};
    const float dH[4] = {-hh * step_gt(t.x, a.x1), -hw * step_gt(t.y, a.y1), hh * step_gt(a.x2, t.z), hw * step_gt(a.y2, t.w)};
    const float ru = __frcp_rn(uni + eps), rh = __frcp_rn(hull + eps);
    const float k_iou = inter * ru * ru, k_pen = (hull - uni) * rh * rh;
    float dxy[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float dU = dA[c] - dI[c];
        const float d_iou = dI[c] * ru - k_iou * dU;
        const float d_pen = (dH[c] - dU) * rh - k_pen * dH[c];
        dxy[c] = s_gi * (d_pen - d_iou);
    }
    dcx += dxy[0] + dxy[2]; dcy += dxy[1] + dxy[3];
    dw += 0.5f * (dxy[2] - dxy[0]); dh += 0.5f * (dxy[3] - dxy[1]);
    return make_float4(dcx, dcy, dw, dh);
}

// One CTA per (image, layer), one thread per query row (as the dense forward kernel): the block is staged by bulk copies,
// every row is turned into its gradient IN PLACE (c * (softmax - onehot): coefficient, log-sum-exp and target class live in
// the owning thread's registers) and the block leaves by bulk stores -- no per-thread global stores, no per-row shared
// arrays, 36.8 KB of shared memory per problem (6 problems resident per SM).
__global__ void __launch_bounds__(kRowThreads) criterion_bwd_dense_kernel(const CritParams p) {
    extern __shared__ __align__(16) int s_dyn[];  // [Q*K] logits -> grad_logits
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    const int b = blockIdx.x / p.L, l = blockIdx.x % p.L;
    const int Q = p.Q, K = p.K;
    float* s_lg = reinterpret_cast<float*>(s_dyn);
    const float* lg = p.logits + b * p.lg_sb + l * p.lg_sl;
    const float* bx = p.boxes + b * p.bx_sb + l * p.bx_sl;
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    __syncthreads();
    if (tid == 0) stage_block(s_lg, lg, (uint32_t)(Q * K) * 4u, &bar);

    const float g_ce = p.grad_losses[l * 5 + 0], g_l1 = p.grad_losses[l * 5 + 2], g_gi = p.grad_losses[l * 5 + 3];
    const float nb = p.num_boxes ? *p.num_boxes : fmaxf((float)p.gt_off[p.B], 1.f);
    const float ce_scale = g_ce * p.w_ce / p.wsum[l];   // d CE / d logits = g * w_ce * w[t]/W * (softmax - onehot)
    const float s_l1 = g_l1 * p.w_l1 / nb, s_gi = g_gi * p.w_giou / nb;
    const float* lse = p.lse + (int64_t)blockIdx.x * Q;
    const int32_t* tgt = p.tgt + (int64_t)blockIdx.x * Q;
    const float4* tbox = reinterpret_cast<const float4*>(p.tbox) + (int64_t)blockIdx.x * Q;
    float4* dbx = reinterpret_cast<float4*>(p.grad_boxes + (int64_t)blockIdx.x * Q * 4);
    bool landed = false;
    const int K4 = K >> 2;
    for (int q = tid; q < Q; q += kRowThreads) {
        // under the copy: coefficients of this row and d boxes (zero for unmatched queries, analytic L1 + GIoU otherwise)
        const float4 t = tbox[q];
        const float4 s = *reinterpret_cast<const float4*>(bx + (int64_t)q * p.bx_sq);
        const int tc_ = tgt[q];
        const float nbias = -lse[q] * kLog2e;
        const float c = ce_scale * p.class_weight[tc_];
        dbx[q] = (t.x == t.x) ? pair_grads_fast(s, t, s_l1, s_gi) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (!landed) { tc::mbar_wait(&bar, 0); landed = true; }
        float4* row = reinterpret_cast<float4*>(s_lg + q * K);
        const float o_t = c * (ex2_approx(fmaf(s_lg[q * K + tc_], kLog2e, nbias)) - 1.f);
#pragma unroll 4
        for (int j = 0; j < K4; ++j) {
            float4 x = row[j];
            x.x = c * ex2_approx(fmaf(x.x, kLog2e, nbias)); x.y = c * ex2_approx(fmaf(x.y, kLog2e, nbias));
            x.z = c * ex2_approx(fmaf(x.z, kLog2e, nbias)); x.w = c * ex2_approx(fmaf(x.w, kLog2e, nbias));
            row[j] = x;
        }
        s_lg[q * K + tc_] = o_t;
    }
    if (!landed) tc::mbar_wait(&bar, 0);
    tc::fence_proxy_async_smem();   // the rows were written through the generic proxy; the bulk store reads through the async one
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)(Q * K) * 4u;
        const uint32_t chunk = ((bytes / 4 + 15u) / 16u) * 16u;
        char* dst = reinterpret_cast<char*>(p.grad_logits + (int64_t)blockIdx.x * Q * K);
        for (uint32_t off = 0; off < bytes; off += chunk)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(dst + off), "r"(tc::smem_u32(reinterpret_cast<char*>(s_lg) + off)), "r"(min(chunk, bytes - off)) : "memory");
        tc::tma_store_commit();
        tc::tma_store_wait_read0();   // shared memory must stay valid until the copy engine has read it
    }
}

static int check_common(const CritParams& p, const char* who) {
    DETR_CHECK_ARG(p.B >= 1 && p.L >= 1 && p.Q >= 1 && p.K >= 1, "%s: bad sizes B=%d L=%d Q=%d K=%d", who, p.B, p.L, p.Q, p.K);
    DETR_CHECK_ARG(((uintptr_t)p.boxes % 16) == 0 && (p.bx_sb % 4) == 0 && (p.bx_sl % 4) == 0 && (p.bx_sq % 4) == 0, "%s: pred boxes must be 16-byte aligned rows", who);
    return 0;
}

}  // namespace detr

using namespace detr;

extern "C" int detr_criterion_fwd_f32(const float* logits, int64_t lg_sb, int64_t lg_sl, int64_t lg_sq,
                                      const float* boxes, int64_t bx_sb, int64_t bx_sl, int64_t bx_sq,
                                      const int64_t* gt_labels, const float* gt_boxes, const int32_t* gt_off,
                                      const int32_t* match_off, const int64_t* idx_q, const int64_t* idx_gt,
                                      const float* class_weight, const float* num_boxes, int B, int L, int Q, int K,
                                      float w_ce, float w_l1, float w_giou, float* partials, float* lse, int32_t* tgt,
                                      float* tbox, float* wsum, float* losses, int32_t* status, void* stream) {
    CritParams p{};
    p.logits = logits; p.lg_sb = lg_sb; p.lg_sl = lg_sl; p.lg_sq = lg_sq;
    p.boxes = boxes; p.bx_sb = bx_sb; p.bx_sl = bx_sl; p.bx_sq = bx_sq;
    p.gt_labels = gt_labels; p.gt_boxes = gt_boxes; p.gt_off = gt_off; p.match_off = match_off;
    p.idx_q = idx_q; p.idx_gt = idx_gt; p.class_weight = class_weight; p.num_boxes = num_boxes;
    p.B = B; p.L = L; p.Q = Q; p.K = K; p.w_ce = w_ce; p.w_l1 = w_l1; p.w_giou = w_giou;
    p.partials = partials; p.lse = lse; p.tgt = tgt; p.tbox = tbox; p.wsum = wsum; p.losses = losses; p.status = status;
    if (check_common(p, "criterion_fwd")) return 1;
    DETR_CHECK_ARG(((uintptr_t)p.gt_boxes % 16) == 0, "criterion_fwd: gt boxes must be 16-byte aligned");
    DETR_CHECK_ARG(partials && lse && tgt && tbox && wsum && losses && status, "criterion_fwd: null output/workspace");
    DETR_CHECK_ARG(((uintptr_t)tbox % 16) == 0, "criterion_fwd: tbox must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = 3 * (size_t)Q * sizeof(int);
    DETR_CHECK_ARG(smem <= 48 * 1024, "criterion_fwd: Q=%d too large (<= 4096)", Q);
    // dense logits blocks with K % 4 == 0: bulk-copy staged, one thread per query row, assignment expanded in the same CTA
    const size_t dense_smem = ((size_t)Q * K + (size_t)Q) * sizeof(int);
    const bool dense = (K % 4) == 0 && lg_sq == K && (lg_sb % 4) == 0 && (lg_sl % 4) == 0 && ((uintptr_t)logits % 16) == 0 &&
                       dense_smem <= 200 * 1024;
    const int chunks = (K + 31) / 32;
    if (dense) {
        if (dense_smem > 48 * 1024 &&
            cudaFuncSetAttribute(criterion_fwd_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dense_smem) != cudaSuccess) {
            set_error("criterion_fwd: cannot reserve %zu B of shared memory", dense_smem);
            return 2;
        }
        criterion_fwd_dense_kernel<<<B * L, kRowThreads, dense_smem, st>>>(p);
    } else {
        criterion_expand_kernel<<<B * L, kExpandThreads, (size_t)Q * sizeof(int), st>>>(p);
        DETR_CHECK_LAUNCH("criterion_expand");
        if (chunks == 1) criterion_fwd_kernel<1><<<B * L, kCritThreads, smem, st>>>(p);
        else if (chunks == 2) criterion_fwd_kernel<2><<<B * L, kCritThreads, smem, st>>>(p);
        else if (chunks == 3) criterion_fwd_kernel<3><<<B * L, kCritThreads, smem, st>>>(p);
        else if (chunks == 4) criterion_fwd_kernel<4><<<B * L, kCritThreads, smem, st>>>(p);
        else criterion_fwd_kernel<0><<<B * L, kCritThreads, smem, st>>>(p);
    }
    DETR_CHECK_LAUNCH("criterion_fwd");
    launch_pdl(criterion_finalize_kernel, dim3(L), dim3(kFinThreads), 0, st, p);
    DETR_CHECK_LAUNCH("criterion_finalize");
    return 0;
}

extern "C" int detr_criterion_bwd_f32(const float* grad_losses, const float* logits, int64_t lg_sb, int64_t lg_sl,
                                      int64_t lg_sq, const float* boxes, int64_t bx_sb, int64_t bx_sl, int64_t bx_sq,
                                      const int32_t* gt_off, const float* class_weight, const float* num_boxes,
                                      const float* lse, const int32_t* tgt, const float* tbox, const float* wsum,
                                      int B, int L, int Q, int K, float w_ce, float w_l1, float w_giou,
                                      float* grad_logits, float* grad_boxes, void* stream) {
    CritParams p{};
    p.logits = logits; p.lg_sb = lg_sb; p.lg_sl = lg_sl; p.lg_sq = lg_sq;
    p.boxes = boxes; p.bx_sb = bx_sb; p.bx_sl = bx_sl; p.bx_sq = bx_sq;
    p.gt_off = gt_off; p.class_weight = class_weight; p.num_boxes = num_boxes;
    p.B = B; p.L = L; p.Q = Q; p.K = K; p.w_ce = w_ce; p.w_l1 = w_l1; p.w_giou = w_giou;
    p.lse = const_cast<float*>(lse); p.tgt = const_cast<int32_t*>(tgt); p.tbox = const_cast<float*>(tbox);
    p.wsum = const_cast<float*>(wsum);
    p.grad_losses = grad_losses; p.grad_logits = grad_logits; p.grad_boxes = grad_boxes;
    if (check_common(p, "criterion_bwd")) return 1;
    DETR_CHECK_ARG(grad_losses && grad_logits && grad_boxes && lse && tgt && tbox && wsum && gt_off, "criterion_bwd: null pointer");
    DETR_CHECK_ARG(((uintptr_t)grad_boxes % 16) == 0 && ((uintptr_t)tbox % 16) == 0, "criterion_bwd: grad_boxes / tbox must be 16-byte aligned");
    const bool vec = (K % 4) == 0 && lg_sq == K && (lg_sb % 4) == 0 && (lg_sl % 4) == 0 && ((uintptr_t)logits % 16) == 0 &&
                     ((uintptr_t)grad_logits % 16) == 0 && 3 * (size_t)Q * sizeof(int) <= 48 * 1024;
    const size_t dense_smem = (size_t)Q * K * sizeof(int);
    if (vec && dense_smem <= 200 * 1024) {
        if (dense_smem > 48 * 1024 &&
            cudaFuncSetAttribute(criterion_bwd_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dense_smem) != cudaSuccess) {
            set_error("criterion_bwd: cannot reserve %zu B of shared memory", dense_smem);
            return 2;
        }
        criterion_bwd_dense_kernel<<<B * L, kRowThreads, dense_smem, (cudaStream_t)stream>>>(p);
    }
    else if (vec) criterion_bwd_kernel<true><<<B * L, kCritThreads, 3 * (size_t)Q * sizeof(int), (cudaStream_t)stream>>>(p);
    else criterion_bwd_kernel<false><<<B * L, kCritThreads, 0, (cudaStream_t)stream>>>(p);
    DETR_CHECK_LAUNCH("criterion_bwd");
    return 0;
}
