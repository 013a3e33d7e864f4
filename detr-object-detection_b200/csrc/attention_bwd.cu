// Flash attention backward for head_dim = 32 (gradient of detr/model.py:317-352), tcgen05 + TMA + TMEM.
//
// Two deterministic kernels (no atomics) built from one template; both recompute, per 128x128 tile,
//     S = Q K^T,  dP~ = dO V^T,  P = exp(S/sqrt(d) - LSE),  P~ = dropout(P),
//     dS = P o (dropout(dP~) - D) / sqrt(d),     D = rowsum(dO o O)
// with queries on the TMEM lanes (so LSE, D and the dropout key are per-thread scalars, exactly as in forward):
//   kDQ   CTA = (batch, head, 128-query tile), streams key tiles :  dQ += dS K            (dS as K-major A operand)
//   kDKV  CTA = (batch, head, 128-key tile),  streams query tiles:  dV += P~^T dO, dK += dS^T Q
//                                                                   (P~, dS consumed as MN-major A operands)
// Warps 0-7 compute (warp w: TMEM lanes 32*(w%4).., key columns 64*(w/4)..), warp 8 = TMA, warp 9 = MMA issue.
// TMEM (512 columns): S [0,128) | dP [128,256) | acc0 [256,288) (dQ or dK) | acc1 [288,320) (dV).
// S/dP are copied to registers and released immediately, so the next tile's score MMAs overlap this tile's math.
//
// Masked keys (key_padding_mask / attention_mask) receive zero gradient as in the reference (masked_fill);
// a query row whose keys are ALL masked is stored with LSE = +inf by the forward kernel and contributes nothing.
#include <math_constants.h>

#include <type_traits>

#include "common.cuh"
#include "tc.cuh"

namespace detr {
using namespace tc;

int make_head_tile_map(CUtensorMap* out, const void* base, int C, int rows, int B, int64_t row_stride_el, int64_t batch_stride_el,
                       int box_rows, const char* who);

namespace bwd {

constexpr int kT = 128;            // tile edge (queries and keys)
constexpr int kD = 32;
constexpr int kStages = 2;
constexpr int kThreads = 320;
constexpr uint32_t kTileBytes = kT * kD * 2;  // 8 KB
constexpr uint32_t kTmemCols = 512;

struct Params {
    // outputs (bf16, channel stride 1)
    __nv_bfloat16* out0; int64_t o0_sb, o0_sl;   // dQ (kDQ)  or dK (kDKV)
    __nv_bfloat16* out1; int64_t o1_sb, o1_sl;   // unused     or dV
    const float* lse;     // (B, nh, L) natural log
    const float* delta;   // (B, nh, L) rowsum(dO o O)
    const uint8_t* kpm; int64_t kpm_sb;
    const uint8_t* amask;
    int B, nh, L, S;
    float scale_log2, scale;   // log2(e)/sqrt(d), 1/sqrt(d)
    uint32_t drop_thresh; float drop_scale; uint64_t seed; const uint64_t* seed_ptr;
};

struct Smem {
    static constexpr uint32_t fixed = 0;                                   // 2 tiles
    static constexpr uint32_t ring = fixed + 2 * kTileBytes;               // kStages x 2 tiles
    static constexpr uint32_t ds = ring + kStages * 2 * kTileBytes;        // 128x128 bf16, SWIZZLE_128B, two 64-key blocks
    static constexpr uint32_t pt = ds + kT * kT * 2;                       // P~ (kDKV only)
    static constexpr uint32_t bars = pt + kT * kT * 2;
    static constexpr uint32_t flags = bars + 128;
    // + ceil(S/128)*128 key flags + 1024 bytes of alignment slack (computed on the host)
};
static_assert(Smem::ds % 1024 == 0 && Smem::pt % 1024 == 0, "swizzled tiles must be 1024-byte aligned");

template <bool kDQ>
__global__ void __launch_bounds__(kThreads, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t0 = blockIdx.x * kT, h = blockIdx.y, b = blockIdx.z;   // fixed tile origin (queries for kDQ, keys for kDKV)
    const int T = kDQ ? (p.S + kT - 1) / kT : (p.L + kT - 1) / kT;    // number of streamed tiles

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
    uint64_t* fixed_full = bars + 0;
    uint64_t* ring_full = bars + 1;
    uint64_t* ring_empty = bars + 1 + kStages;
    uint64_t* sdp_full = bars + 1 + 2 * kStages;
    uint64_t* sdp_empty = sdp_full + 1;
    uint64_t* ds_full = sdp_full + 2;
    uint64_t* ds_empty = sdp_full + 3;
    uint64_t* acc_full = sdp_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sdp_full + 5);
    uint8_t* kflag = smem + Smem::flags;

    if (tid == 0) {
        mbar_init(fixed_full, 1);
        for (int s = 0; s < kStages; ++s) { mbar_init(ring_full + s, 1); mbar_init(ring_empty + s, 1); }
        mbar_init(sdp_full, 1); mbar_init(sdp_empty, 256); mbar_init(ds_full, 256); mbar_init(ds_empty, 1); mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_slot, kTmemCols);
    {   // key flags: 0 normal, 1 masked, 2 beyond S.  kDQ: all keys; kDKV: the CTA's own 128 keys
        const int nk = kDQ ? T * kT : kT, base = kDQ ? 0 : t0;
        for (int k = tid; k < nk; k += kThreads)
            kflag[k] = (base + k) >= p.S ? 2 : ((p.kpm && p.kpm[b * p.kpm_sb + base + k]) ? 1 : 0);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 128, tmem_a0 = tmem_base + 256, tmem_a1 = tmem_base + 288;

    if (warp == 8) {
        // ================= TMA producer =================
        if (lane == 0) {
            tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_do);
            mbar_expect_tx(fixed_full, 2 * kTileBytes);
            tma_load_3d(smem + Smem::fixed, kDQ ? &tm_q : &tm_k, fixed_full, h * kD, t0, b);
            tma_load_3d(smem + Smem::fixed + kTileBytes, kDQ ? &tm_do : &tm_v, fixed_full, h * kD, t0, b);
            for (int t = 0; t < T; ++t) {
                const int s = t % kStages;
                if (t >= kStages) mbar_wait_sleep(ring_empty + s, ((t / kStages) - 1) & 1);
                mbar_expect_tx(ring_full + s, 2 * kTileBytes);
                uint8_t* dst = smem + Smem::ring + s * 2 * kTileBytes;
                tma_load_3d(dst, kDQ ? &tm_k : &tm_q, ring_full + s, h * kD, t * kT, b);
                tma_load_3d(dst + kTileBytes, kDQ ? &tm_v : &tm_do, ring_full + s, h * kD, t * kT, b);
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc_sc = make_idesc_bf16(kT, kT, false, false);   // scores: both operands K-major (contract over d)
            constexpr uint32_t idesc_dq = make_idesc_bf16(kT, kD, false, true);    // dS (K-major) x K (MN-major)
            constexpr uint32_t idesc_kv = make_idesc_bf16(kT, kD, true, true);     // P~^T / dS^T (MN-major) x dO / Q (MN-major)
            const uint32_t fx0 = smem_u32(smem + Smem::fixed), fx1 = fx0 + kTileBytes;
            const uint32_t sds = smem_u32(smem + Smem::ds), spt = smem_u32(smem + Smem::pt);
            auto kmaj64 = [](uint32_t a, int ks) { return make_smem_desc(a + ks * 32, 16, 512, SWZ_64B); };
            auto issue_scores = [&](int t) {
                const uint32_t r0 = smem_u32(smem + Smem::ring + (t % kStages) * 2 * kTileBytes), r1 = r0 + kTileBytes;
                const uint32_t aq = kDQ ? fx0 : r0, bk = kDQ ? r0 : fx0, ado = kDQ ? fx1 : r1, bv = kDQ ? r1 : fx1;
#pragma unroll
                for (int ks = 0; ks < kD / 16; ++ks) umma_bf16(tmem_s, kmaj64(aq, ks), kmaj64(bk, ks), idesc_sc, ks > 0);
#pragma unroll
                for (int ks = 0; ks < kD / 16; ++ks) umma_bf16(tmem_dp, kmaj64(ado, ks), kmaj64(bv, ks), idesc_sc, ks > 0);
                umma_commit(sdp_full);
            };
            mbar_wait_sleep(fixed_full, 0);
            mbar_wait_sleep(ring_full + 0, 0);
            tc_fence_after();
            issue_scores(0);
            for (int t = 0; t < T; ++t) {
                if (t + 1 < T) {
                    mbar_wait_sleep(ring_full + ((t + 1) % kStages), ((t + 1) / kStages) & 1);
                    mbar_wait_sleep(sdp_empty, t & 1);
                    tc_fence_after();
                    issue_scores(t + 1);
                }
                mbar_wait_sleep(ds_full, t & 1);
                tc_fence_after();
                const uint32_t r0 = smem_u32(smem + Smem::ring + (t % kStages) * 2 * kTileBytes), r1 = r0 + kTileBytes;
#pragma unroll
                for (int ks = 0; ks < kT / 16; ++ks) {
                    if (kDQ) {
                        // dQ += dS K_t : A K-major SWIZZLE_128B (64-key blocks of 16 KB), B = K tile MN-major SWIZZLE_64B
                        umma_bf16(tmem_a0, make_smem_desc(sds + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, SWZ_128B),
                                  make_smem_desc(r0 + ks * 1024, 512, 512, SWZ_64B), idesc_dq, t > 0 || ks > 0);
                    } else {
                        // contraction over queries: A = [query][key] tiles read MN-major (keys = M): 16 queries = 2048 B per step,
                        // the second 64-key block 16 KB further (LBO), 8-query groups 1024 B apart (SBO)
                        const uint64_t b_do = make_smem_desc(r1 + ks * 1024, 512, 512, SWZ_64B);
                        const uint64_t b_q = make_smem_desc(r0 + ks * 1024, 512, 512, SWZ_64B);
                        umma_bf16(tmem_a1, make_smem_desc(spt + ks * 2048, 16384, 1024, SWZ_128B), b_do, idesc_kv, t > 0 || ks > 0);
                        umma_bf16(tmem_a0, make_smem_desc(sds + ks * 2048, 16384, 1024, SWZ_128B), b_q, idesc_kv, t > 0 || ks > 0);
                    }
                }
                umma_commit(ds_empty);
                umma_commit(ring_empty + (t % kStages));
            }
            umma_commit(acc_full);
        }
    } else {
        // ================= compute warps =================
        const int row = (warp & 3) * 32 + lane;       // TMEM lane = query within the tile
        const int ch = warp >> 2;                     // which 64-key half of the tile
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t bh = (uint32_t)(b * p.nh + h);
        const float* lse_bh = p.lse + (int64_t)bh * p.L;
        const float* dl_bh = p.delta + (int64_t)bh * p.L;
        uint8_t* ds_row = smem + Smem::ds + ch * 16384 + (row >> 3) * 1024 + (row & 7) * 128;
        uint8_t* pt_row = smem + Smem::pt + ch * 16384 + (row >> 3) * 1024 + (row & 7) * 128;
        const bool drop = p.drop_thresh != 0;

        float lse2 = 0.f, dlt = 0.f;
        uint32_t row_key = 0;
        int q = 0;
        auto load_row = [&](int qq) {
            q = qq;
            const bool ok = q < p.L;
            const float l = ok ? lse_bh[q] : CUDART_INF_F;
            lse2 = l * 1.4426950408889634f;            // +inf (padding row / fully masked row) -> p = 0
            dlt = ok ? dl_bh[q] : 0.f;
            row_key = drop ? dropout_row_key(p.seed + (p.seed_ptr ? *p.seed_ptr : 0ull), bh, (uint32_t)q) : 0u;
        };
        if (kDQ) load_row(t0 + row);

        uint32_t s0[32], s1[32], d0[32], d1[32];
        for (int t = 0; t < T; ++t) {
            const int key0 = (kDQ ? t * kT : t0) + ch * 64;      // first key of this thread's 64-column half
            if (!kDQ) load_row(t * kT + row);
            const uint8_t* kf = kflag + (kDQ ? t * kT : 0) + ch * 64;
            const uint8_t* arow = (p.amask && q < p.L) ? p.amask + (int64_t)q * p.S + key0 : nullptr;
            uint32_t any = p.amask ? 1u : 0u;
            {
                const uint4* kf4 = reinterpret_cast<const uint4*>(kf);
#pragma unroll
                for (int i = 0; i < 4; ++i) { const uint4 w = kf4[i]; any |= w.x | w.y | w.z | w.w; }
            }
            mbar_wait(sdp_full, t & 1);
            tc_fence_after();
            tmem_ld32(tmem_s + lane_addr + ch * 64, s0);
            tmem_ld32(tmem_s + lane_addr + ch * 64 + 32, s1);
            tmem_ld32(tmem_dp + lane_addr + ch * 64, d0);
            tmem_ld32(tmem_dp + lane_addr + ch * 64 + 32, d1);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(sdp_empty);                         // score columns may be overwritten by the next tile
            if (t > 0) mbar_wait(ds_empty, (t - 1) & 1);    // previous accumulation MMAs have consumed dS / P~

            auto half_tile = [&](auto masked_c, auto drop_c, const uint32_t* sv, const uint32_t* dv, int hf) {
                constexpr bool MASKED = decltype(masked_c)::value, DROP = decltype(drop_c)::value;
                const uint32_t th = p.drop_thresh << 24;
#pragma unroll
                for (int g = 0; g < 4; ++g) {               // 8 keys = one 16-byte chunk of the dS / P~ rows
                    float pt[8], dsv[8];
                    uint32_t b0 = 0, b1 = 0;
                    if (DROP) {
                        const uint32_t k4 = (uint32_t)((key0 + hf * 32 + g * 8) >> 2);
                        b0 = dropout_bits4(row_key, k4); b1 = dropout_bits4(row_key, k4 + 1);
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int i = g * 8 + e;
                        float pr = ex2(fmaf(__uint_as_float(sv[i]), p.scale_log2, -lse2));
                        if (MASKED) {
                            const int c = hf * 32 + i;
                            const bool m = kf[c] != 0 || (arow != nullptr && (key0 + c) < p.S && arow[c] != 0);
                            pr = m ? 0.f : pr;
                        }
                        float keep = 1.f;
                        if (DROP) keep = dropout_keep(e < 4 ? b0 : b1, e & 3, th) ? p.drop_scale : 0.f;
                        pt[e] = DROP ? pr * keep : pr;
                        // dS without the 1/sqrt(d) factor: it is applied once to the dQ / dK accumulators in the epilogue
                        dsv[e] = pr * (DROP ? fmaf(__uint_as_float(dv[i]), keep, -dlt) : (__uint_as_float(dv[i]) - dlt));
                    }
                    const uint32_t off = (uint32_t)(((hf * 4 + g) ^ (row & 7)) << 4);
                    uint4 w;
                    w.x = pack_bf16x2(dsv[0], dsv[1]); w.y = pack_bf16x2(dsv[2], dsv[3]);
                    w.z = pack_bf16x2(dsv[4], dsv[5]); w.w = pack_bf16x2(dsv[6], dsv[7]);
                    *reinterpret_cast<uint4*>(ds_row + off) = w;
                    if (!kDQ) {
                        w.x = pack_bf16x2(pt[0], pt[1]); w.y = pack_bf16x2(pt[2], pt[3]);
                        w.z = pack_bf16x2(pt[4], pt[5]); w.w = pack_bf16x2(pt[6], pt[7]);
                        *reinterpret_cast<uint4*>(pt_row + off) = w;
                    }
                }
            };
            using TT = std::true_type; using FF = std::false_type;
            if (any) {
                if (drop) { half_tile(TT{}, TT{}, s0, d0, 0); half_tile(TT{}, TT{}, s1, d1, 1); }
                else      { half_tile(TT{}, FF{}, s0, d0, 0); half_tile(TT{}, FF{}, s1, d1, 1); }
            } else {
                if (drop) { half_tile(FF{}, TT{}, s0, d0, 0); half_tile(FF{}, TT{}, s1, d1, 1); }
                else      { half_tile(FF{}, FF{}, s0, d0, 0); half_tile(FF{}, FF{}, s1, d1, 1); }
            }
            fence_proxy_async_smem();
            mbar_arrive(ds_full);
        }
        // ---- epilogue: accumulators -> bf16 global ----
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int n_rows = kDQ ? p.L : p.S;
        const int gr = t0 + row;
        if (kDQ ? (ch == 0) : true) {
            tmem_ld32((kDQ || ch == 0 ? tmem_a0 : tmem_a1) + lane_addr, s0);
            tmem_ld_wait();
            if (gr < n_rows) {
                __nv_bfloat16* dst = (kDQ || ch == 0) ? p.out0 + b * p.o0_sb + (int64_t)gr * p.o0_sl + h * kD
                                                      : p.out1 + b * p.o1_sb + (int64_t)gr * p.o1_sl + h * kD;
                const float f = (kDQ || ch == 0) ? p.scale : 1.f;   // dQ and dK carry the 1/sqrt(d) of the scores; dV does not
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 w;
                    w.x = pack_bf16x2(f * __uint_as_float(s0[g * 8 + 0]), f * __uint_as_float(s0[g * 8 + 1]));
                    w.y = pack_bf16x2(f * __uint_as_float(s0[g * 8 + 2]), f * __uint_as_float(s0[g * 8 + 3]));
                    w.z = pack_bf16x2(f * __uint_as_float(s0[g * 8 + 4]), f * __uint_as_float(s0[g * 8 + 5]));
                    w.w = pack_bf16x2(f * __uint_as_float(s0[g * 8 + 6]), f * __uint_as_float(s0[g * 8 + 7]));
                    reinterpret_cast<uint4*>(dst)[g] = w;
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, kTmemCols);
}

// D[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]   (one thread per (b,q,h), 64-byte vector loads)
__global__ void attention_delta_kernel(const __nv_bfloat16* __restrict__ dO, int64_t do_sb, int64_t do_sl,
                                       const __nv_bfloat16* __restrict__ O, int64_t o_sb, int64_t o_sl,
                                       float* __restrict__ delta, int B, int nh, int L) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)B * L * nh) return;
    const int h = (int)(idx % nh);
    const int64_t bq = idx / nh;
    const int q = (int)(bq % L), b = (int)(bq / L);
    const uint4* a = reinterpret_cast<const uint4*>(dO + b * do_sb + (int64_t)q * do_sl + h * 32);
    const uint4* c = reinterpret_cast<const uint4*>(O + b * o_sb + (int64_t)q * o_sl + h * 32);
    float acc = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint4 x = a[g], y = c[g];
        const __nv_bfloat162* xv = reinterpret_cast<const __nv_bfloat162*>(&x);
        const __nv_bfloat162* yv = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 fx = __bfloat1622float2(xv[e]), fy = __bfloat1622float2(yv[e]);
            acc = fmaf(fx.x, fy.x, acc);
            acc = fmaf(fx.y, fy.y, acc);
        }
    }
    delta[((int64_t)b * nh + h) * L + q] = acc;
}

}  // namespace bwd
}  // namespace detr

using namespace detr;

extern "C" int detr_attention_bwd_bf16(const void* q, int64_t q_sb, int64_t q_sl, const void* k, int64_t k_sb, int64_t k_sl,
                                       const void* v, int64_t v_sb, int64_t v_sl, const void* o, int64_t o_sb, int64_t o_sl,
                                       const void* d_o, int64_t do_sb, int64_t do_sl, const float* lse, float* delta,
                                       void* dq, int64_t dq_sb, int64_t dq_sl, void* dk, int64_t dk_sb, int64_t dk_sl,
                                       void* dv, int64_t dv_sb, int64_t dv_sl, const uint8_t* key_padding_mask, int64_t kpm_sb,
                                       const uint8_t* attention_mask, int B, int nh, int L, int S, float dropout_p,
                                       uint64_t seed, const uint64_t* seed_ptr, void* stream) {
    using namespace detr::bwd;
    DETR_CHECK_ARG(B >= 1 && nh >= 1 && L >= 1 && S >= 1, "attention_bwd: bad sizes B=%d nh=%d L=%d S=%d", B, nh, L, S);
    DETR_CHECK_ARG(B <= 65535 && nh <= 65535, "attention_bwd: B and nh must fit the grid");
    DETR_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "attention_bwd: dropout_p must be in [0,1)");
    auto aligned = [](const void* ptr, int64_t sb, int64_t sl) { return ((uintptr_t)ptr % 16) == 0 && (sb % 8) == 0 && (sl % 8) == 0; };
    DETR_CHECK_ARG(aligned(dq, dq_sb, dq_sl) && aligned(dk, dk_sb, dk_sl) && aligned(dv, dv_sb, dv_sl) && aligned(o, o_sb, o_sl) &&
                       aligned(d_o, do_sb, do_sl),
                   "attention_bwd: tensors must be 16-byte aligned with row/batch strides multiple of 8 elements");
    cudaStream_t st = (cudaStream_t)stream;
    const int C = nh * kD;
    CUtensorMap tq, tk, tv, tdo;
    if (int rc = make_head_tile_map(&tq, q, C, L, B, q_sl, q_sb, kT, "attention_bwd(Q)")) return rc;
    if (int rc = make_head_tile_map(&tk, k, C, S, B, k_sl, k_sb, kT, "attention_bwd(K)")) return rc;
    if (int rc = make_head_tile_map(&tv, v, C, S, B, v_sl, v_sb, kT, "attention_bwd(V)")) return rc;
    if (int rc = make_head_tile_map(&tdo, d_o, C, L, B, do_sl, do_sb, kT, "attention_bwd(dO)")) return rc;

    const int64_t n = (int64_t)B * L * nh;
    attention_delta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(d_o), do_sb, do_sl, reinterpret_cast<const __nv_bfloat16*>(o), o_sb, o_sl, delta, B, nh, L);
    DETR_CHECK_LAUNCH("attention_delta");

    Params p;
    p.lse = lse; p.delta = delta; p.kpm = key_padding_mask; p.kpm_sb = kpm_sb; p.amask = attention_mask;
    p.B = B; p.nh = nh; p.L = L; p.S = S;
    p.scale = 1.f / sqrtf((float)kD);
    p.scale_log2 = 1.4426950408889634f * p.scale;
    p.drop_thresh = (uint32_t)lrintf(dropout_p * 256.f);
    p.drop_scale = 256.f / (256.f - (float)p.drop_thresh);
    p.seed = seed; p.seed_ptr = seed_ptr;
    const size_t smem_dq = Smem::flags + (size_t)((S + kT - 1) / kT) * kT + 1024, smem_dkv = Smem::flags + kT + 1024;
    DETR_CHECK_ARG(smem_dq <= 200 * 1024, "attention_bwd: S=%d needs %zu B of shared memory", S, smem_dq);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e1 = cudaFuncSetAttribute(attention_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaError_t e2 = cudaFuncSetAttribute(attention_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("attention_bwd: cudaFuncSetAttribute failed"); return 2; }
        attr_set = true;
    }
    // dK, dV
    p.out0 = reinterpret_cast<__nv_bfloat16*>(dk); p.o0_sb = dk_sb; p.o0_sl = dk_sl;
    p.out1 = reinterpret_cast<__nv_bfloat16*>(dv); p.o1_sb = dv_sb; p.o1_sl = dv_sl;
    attention_bwd_kernel<false><<<dim3((S + kT - 1) / kT, nh, B), kThreads, smem_dkv, st>>>(tq, tk, tv, tdo, p);
    DETR_CHECK_LAUNCH("attention_bwd_dkv");
    // dQ
    p.out0 = reinterpret_cast<__nv_bfloat16*>(dq); p.o0_sb = dq_sb; p.o0_sl = dq_sl;
    p.out1 = nullptr; p.o1_sb = p.o1_sl = 0;
    attention_bwd_kernel<true><<<dim3((L + kT - 1) / kT, nh, B), kThreads, smem_dq, st>>>(tq, tk, tv, tdo, p);
    DETR_CHECK_LAUNCH("attention_bwd_dq");
    return 0;
}
