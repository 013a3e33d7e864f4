// Flash attention backward for head_dim = 32 (gradient of detr/model.py:317-352), tcgen05 + TMA + TMEM.
//
// ONE pass over the (query tile, key tile) pairs.  CTA = (batch, head, 128-key tile); it streams the query tiles and,
// per 128x128 pair, recomputes
//     S = Q K^T,  dP~ = dO V^T,  P = exp(S/sqrt(d) - LSE),  P~ = dropout(P),
//     dS = P o (dropout(dP~) - D) / sqrt(d),     D = rowsum(dO o O)
// exactly once, with queries on the TMEM lanes (so LSE, D and the dropout key are per-thread scalars, as in forward):
//     dV += P~^T dO,  dK += dS^T Q      accumulate in TMEM over the whole stream (P~, dS consumed as MN-major A operands)
//     dQ_part = dS K                     per pair (the SAME dS tile consumed as a K-major A operand), written as an fp32
//                                        partial per key tile; attention_dq_reduce_kernel sums the partials in a fixed
//                                        order -> deterministic, no atomics.
// 20 warps: 0-15 compute (warp w: TMEM lanes 32*(w%4).., key columns 32*(w/4)..+31 of the pair), 16 = TMA producer,
// 17 = tcgen05.mma issuer, 18 = dQ-partial store warp (TMA bulk tensor stores of tiles the math warps stage in shared
// memory: per-row 32-byte global stores from 512 threads cost ~600 cycles per pair in the LSU), 19 idle (completes the
// fifth warpgroup so that setmaxnreg can move its registers to the math).
// TMEM (512 columns): S [0,128) | dP [128,256) | dK [256,288) | dV [288,320) | dQ_part x3 [320,416).
// S / dP are copied to registers and released at once, so the next pair's score MMAs run under this pair's math;
// the dS / P~ shared-memory tiles and the dQ_part columns are double buffered, so the math never waits for the MMAs.
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
int make_f32_tile_map(CUtensorMap* out, const void* base, int C, int rows, int slabs, int box_c, int box_rows, const char* who);

namespace bwd {

constexpr int kT = 128;            // tile edge (queries and keys)
constexpr int kD = 32;
constexpr int kStages = 3;
constexpr int kComputeWarps = 16;
constexpr int kComputeThreads = kComputeWarps * 32;
constexpr int kThreads = 640;      // 5 warpgroups
constexpr uint32_t kTileBytes = kT * kD * 2;  // 8 KB
constexpr uint32_t kTmemCols = 512;

struct Params {
    __nv_bfloat16* dk; int64_t dk_sb, dk_sl;
    __nv_bfloat16* dv; int64_t dv_sb, dv_sl;
    float* dq_part;       // [key tiles][B][L][nh*32] fp32 partial dQ (without the 1/sqrt(d) factor)
    const float* lse;     // (B, nh, L) natural log
    const float* delta;   // (B, nh, L) rowsum(dO o O)
    const uint8_t* kpm; int64_t kpm_sb;
    const uint8_t* amask;
    int B, nh, L, S;
    float scale_log2, scale;   // log2(e)/sqrt(d), 1/sqrt(d)
    uint32_t drop_thresh;      // 0 = no dropout; a key is dropped if its 7 random bits < thresh
    float drop_log2_scale;     // log2(128 / (128 - thresh)): folded into the exponent, P comes out pre-scaled by 1/(1-p)
    float drop_keep;           // (128 - thresh) / 128
    uint64_t seed; const uint64_t* seed_ptr;
    long long* dbg;            // optional timeline of CTA 0 (clock64 stamps; tools/attn_timeline.py), NULL in production
};

struct Smem {
    static constexpr uint32_t fixed = 0;                                   // K tile, V tile
    static constexpr uint32_t ring = fixed + 2 * kTileBytes;               // kStages x (Q tile, dO tile)
    static constexpr uint32_t ds = 65536;                                  // 2 x (128x128 bf16, SWIZZLE_128B, two 64-key blocks)
    static constexpr uint32_t pt = ds + 2 * kT * kT * 2;                   // 2 x P~
    static constexpr uint32_t dqs = pt + 2 * kT * kT * 2;                  // 2 x (128 queries x 32 fp32, SWIZZLE_128B) dQ_part staging
    static constexpr uint32_t bars = dqs + 2 * kT * kD * 4;
    static constexpr uint32_t flags = bars + 256;                          // 128 key flags
    static constexpr uint32_t total = flags + kT + 1024;                   // + alignment slack
};
static_assert(Smem::ring + kStages * 2 * kTileBytes <= Smem::ds && Smem::ds % 1024 == 0 && Smem::pt % 1024 == 0, "smem layout");

// debug timeline (build with DETR_B200_DEFINES=-DDETR_BWD_TIMELINE): slot = (warp, tile, event) of CTA (0,0,0) only
#ifndef DETR_BWD_TIMELINE
#define BWD_STAMP(ev, t) do {} while (0)
#else
#define BWD_STAMP(ev, t) do { if (p.dbg != nullptr && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (t) < 16) \
    p.dbg[(warp * 16 + (t)) * 8 + (ev)] = clock64(); } while (0)
#endif

__global__ void __launch_bounds__(kThreads, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                     const __grid_constant__ CUtensorMap tm_dqp, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kt = blockIdx.x, k0 = kt * kT, h = blockIdx.y, b = blockIdx.z;
#ifdef DETR_BWD_TIMELINE
    const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (p.dbg != nullptr && tid == 0 && cta_lin < 1024) {
        uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.dbg[2560 + cta_lin * 4 + 0] = gt; p.dbg[2560 + cta_lin * 4 + 3] = smid;
        if (cta_lin == 0) p.dbg[(19 * 16 + 0) * 8 + 0] = clock64();
    }
#endif
    const int T = (p.L + kT - 1) / kT;    // number of streamed query tiles

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
    uint64_t* fixed_full = bars + 0;
    uint64_t* ring_full = bars + 1;                    // [kStages]
    uint64_t* ring_empty = ring_full + kStages;        // [kStages]
    uint64_t* sdp_full = ring_empty + kStages;         // [2] one per 64-key half: the two math groups run out of phase
    uint64_t* sdp_empty = sdp_full + 2;                // [2]
    uint64_t* ds_full = sdp_full + 4;                  // [2]
    uint64_t* dq_full = sdp_full + 6;                  // [3] dQ_part TMEM columns are triple buffered: read out two pairs late
    uint64_t* dq_empty = sdp_full + 9;                 // [3]
    uint64_t* dqs_full = sdp_full + 12;                // [2]
    uint64_t* dqs_empty = sdp_full + 14;               // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sdp_full + 16);
    uint8_t* kflag = smem + Smem::flags;

    if (tid == 0) {
        mbar_init(fixed_full, 1);
        for (int s = 0; s < kStages; ++s) { mbar_init(ring_full + s, 1); mbar_init(ring_empty + s, 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(sdp_full + s, 1); mbar_init(sdp_empty + s, kComputeThreads / 2);
            mbar_init(ds_full + s, kComputeThreads);
            mbar_init(dqs_full + s, kComputeThreads); mbar_init(dqs_empty + s, 1);
        }
        for (int s = 0; s < 3; ++s) { mbar_init(dq_full + s, 1); mbar_init(dq_empty + s, kComputeThreads); }
        fence_barrier_init();
    }
    if (warp == 17) tmem_alloc(tmem_slot, kTmemCols);
    if (tid < kT)   // key flags of the CTA's own 128 keys: 0 normal, 1 masked, 2 beyond S
        kflag[tid] = (k0 + tid) >= p.S ? 2 : ((p.kpm && p.kpm[b * p.kpm_sb + k0 + tid]) ? 1 : 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
#ifdef DETR_BWD_TIMELINE
    if (p.dbg != nullptr && tid == 0 && cta_lin < 1024) {
        long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.dbg[2560 + cta_lin * 4 + 1] = gt;
    }
#endif
    const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 128, tmem_dk = tmem_base + 256, tmem_dv = tmem_base + 288,
                   tmem_dq = tmem_base + 320;

    if (warp >= kComputeWarps) {
        reg_dealloc<64>();   // the CTA register pool is what its own warps release: 128 x (96-64) == 512 x (104-96)
        if (warp == 16 && lane == 0) {
            // ================= TMA producer =================
            tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_do);
            mbar_expect_tx(fixed_full, 2 * kTileBytes);
            tma_load_3d(smem + Smem::fixed, &tm_k, fixed_full, h * kD, k0, b);
            tma_load_3d(smem + Smem::fixed + kTileBytes, &tm_v, fixed_full, h * kD, k0, b);
            for (int t = 0; t < T; ++t) {
                const int s = t % kStages;
                if (t >= kStages) mbar_wait_sleep(ring_empty + s, ((t / kStages) - 1) & 1);
                mbar_expect_tx(ring_full + s, 2 * kTileBytes);
                uint8_t* dst = smem + Smem::ring + s * 2 * kTileBytes;
                tma_load_3d(dst, &tm_q, ring_full + s, h * kD, t * kT, b);
                tma_load_3d(dst + kTileBytes, &tm_do, ring_full + s, h * kD, t * kT, b);
            }
        } else if (warp == 17 && elect_one()) {
            // ================= MMA issuer =================
            constexpr uint32_t idesc_sc = make_idesc_bf16(kT, kT, false, false);   // scores: both operands K-major (contract over d)
            constexpr uint32_t idesc_dq = make_idesc_bf16(kT, kD, false, true);    // dS (K-major) x K (MN-major)
            constexpr uint32_t idesc_kv = make_idesc_bf16(kT, kD, true, true);     // P~^T / dS^T (MN-major) x dO / Q (MN-major)
            // descriptor words (tc.cuh): high word per layout, low word = (address >> 4) + constant
            constexpr uint32_t hi64 = desc_hi(512, SWZ_64B);      // Q/K/V/dO tiles: 64-byte rows, 8-row groups 512 B apart
            constexpr uint32_t hi128 = desc_hi(1024, SWZ_128B);   // dS / P~ tiles: 128-byte rows, 8-row groups 1024 B apart
            const uint32_t k_lo = smem_u32(smem + Smem::fixed) >> 4, v_lo = k_lo + (kTileBytes >> 4);
            constexpr uint32_t idesc_half = make_idesc_bf16(kT, kT / 2, false, false);
            // scores of one 64-key half: S_h = Q K_h^T, dP_h = dO V_h^T (keys 64h.. of the K / V tiles: 64 rows x 64 B further)
            auto issue_scores = [&](int t, int hf) {
                const uint32_t q_lo = smem_u32(smem + Smem::ring + (t % kStages) * 2 * kTileBytes) >> 4, do_lo = q_lo + (kTileBytes >> 4);
                const uint32_t kh_lo = k_lo + hf * (64 * 64 >> 4), vh_lo = v_lo + hf * (64 * 64 >> 4);
#pragma unroll
                for (int ks = 0; ks < kD / 16; ++ks)   // K-major operands: 32 B per 16-channel step, LBO 16
                    umma_bf16_lh(tmem_s + hf * 64, q_lo + desc_lo(ks * 32, 16), hi64, kh_lo + desc_lo(ks * 32, 16), hi64, idesc_half, ks > 0);
#pragma unroll
                for (int ks = 0; ks < kD / 16; ++ks)
                    umma_bf16_lh(tmem_dp + hf * 64, do_lo + desc_lo(ks * 32, 16), hi64, vh_lo + desc_lo(ks * 32, 16), hi64, idesc_half, ks > 0);
                umma_commit(sdp_full + hf);
            };
            mbar_wait_sleep(fixed_full, 0);
            mbar_wait_sleep(ring_full + 0, 0);
            tc_fence_after();
            issue_scores(0, 0);
            issue_scores(0, 1);
            for (int t = 0; t < T; ++t) {
                const int buf = t & 1;
                if (t + 1 < T) {
                    mbar_wait_sleep(ring_full + ((t + 1) % kStages), ((t + 1) / kStages) & 1);
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        mbar_wait_sleep(sdp_empty + hf, t & 1);     // that group holds its half of S_t / dP_t in registers
                        tc_fence_after();
                        issue_scores(t + 1, hf);
                    }
                }
                BWD_STAMP(0, t);
                mbar_wait_sleep(ds_full + buf, (t >> 1) & 1);   // dS_t and P~_t are in shared memory
                BWD_STAMP(1, t);
                if (t >= 3) mbar_wait_sleep(dq_empty + t % 3, ((t / 3) - 1) & 1);   // dQ_part of pair t-3 has been read out of TMEM
                if (t >= 2) mbar_wait_sleep(dqs_empty + buf, ((t >> 1) - 1) & 1);   // staging tile of pair t-2 has been stored: dq_full(t)
                                                                                    // tells the math warps that it is free again
                tc_fence_after();
                const uint32_t q_lo = smem_u32(smem + Smem::ring + (t % kStages) * 2 * kTileBytes) >> 4, do_lo = q_lo + (kTileBytes >> 4);
                const uint32_t ds_lo = smem_u32(smem + Smem::ds + buf * (kT * kT * 2)) >> 4, pt_lo = smem_u32(smem + Smem::pt + buf * (kT * kT * 2)) >> 4;
                const bool acc = t > 0;
#pragma unroll
                for (int ks = 0; ks < kT / 16; ++ks) {
                    // contraction over queries: A = [query][key] tiles read MN-major (keys = M): 16 queries = 2048 B per step,
                    // the second 64-key block 16 KB further (LBO), 8-query groups 1024 B apart (SBO);
                    // B = dO / Q tiles MN-major: 16 queries = 1024 B per step, LBO 512
                    umma_bf16_lh(tmem_dv, pt_lo + desc_lo(ks * 2048, 16384), hi128, do_lo + desc_lo(ks * 1024, 512), hi64, idesc_kv, acc || ks > 0);
                    umma_bf16_lh(tmem_dk, ds_lo + desc_lo(ks * 2048, 16384), hi128, q_lo + desc_lo(ks * 1024, 512), hi64, idesc_kv, acc || ks > 0);
                }
#pragma unroll
                for (int ks = 0; ks < kT / 16; ++ks) {
                    // dQ_part = dS K: A K-major SWIZZLE_128B (64-key blocks of 16 KB, 32 B per 16-key step), B = K tile MN-major
                    umma_bf16_lh(tmem_dq + (t % 3) * kD, ds_lo + desc_lo((ks >> 2) * 16384 + (ks & 3) * 32, 16), hi128,
                                 k_lo + desc_lo(ks * 1024, 512), hi64, idesc_dq, ks > 0);
                }
                umma_commit(ring_empty + (t % kStages));
                umma_commit(dq_full + t % 3);                   // (also covers dK / dV of the last pair for the epilogue)
                BWD_STAMP(2, t);
            }
        }
        else if (warp == 18 && elect_one()) {
            // ================= dQ_part store warp =================
            tma_prefetch_desc(&tm_dqp);
            for (int t = 0; t < T; ++t) {
                const int buf = t & 1;
                mbar_wait_sleep(dqs_full + buf, (t >> 1) & 1);
                tma_store_3d(&tm_dqp, smem + Smem::dqs + buf * (kT * kD * 4), h * kD, t * kT, kt * p.B + b);   // rows >= L are clipped
                tma_store_commit();
                tma_store_wait_read0();                 // the staging tile has been read: it may be overwritten
                mbar_arrive(dqs_empty + buf);
            }
            tma_store_wait_all0();                      // the partials are in global memory before the CTA exits
        }
    } else {
        // ================= compute warps =================
        reg_alloc<104>();
        const int lq = warp & 3, kq = warp >> 2;      // TMEM lane quarter, 32-key quarter of the tile
        const int row = lq * 32 + lane;               // TMEM lane = query within the tile
        const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
        const uint32_t bh = (uint32_t)(b * p.nh + h);
        const float* lse_bh = p.lse + (int64_t)bh * p.L;
        const float* dl_bh = p.delta + (int64_t)bh * p.L;
        const bool drop = p.drop_thresh != 0;
        const uint64_t seed = drop ? p.seed + (p.seed_ptr ? *p.seed_ptr : 0ull) : 0ull;
        const uint32_t thr4 = p.drop_thresh * 0x01010101u;
        const int key0 = k0 + kq * 32;                // first key of this thread's 32 columns
        const uint8_t* kf = kflag + kq * 32;
        // byte offset of this thread's row inside a [query][64-key block] SWIZZLE_128B tile; its 4 chunks are (kq&1)*4 + g
        const uint32_t row_off = (uint32_t)((kq >> 1) * 16384 + (row >> 3) * 1024 + (row & 7) * 128);
        const uint32_t chunk0 = (uint32_t)((kq & 1) * 4);

        uint32_t any = p.amask ? 1u : 0u;             // warp-uniform: does this 32-key quarter need masking at all?
        {
            const uint4* kf4 = reinterpret_cast<const uint4*>(kf);
#pragma unroll
            for (int i = 0; i < 2; ++i) { const uint4 w = kf4[i]; any |= w.x | w.y | w.z | w.w; }
        }

        // dQ_part of pair t: 8 of its 32 columns per thread, TMEM -> registers -> swizzled staging tile (row = 128 B,
        // 16-byte chunk j of row r at ((j ^ (r & 7)) << 4): conflict-free stores, the layout TMA expects for SWIZZLE_128B)
        uint8_t* dqs_row = smem + Smem::dqs + row * 128;
        const uint32_t dqs_c0 = (uint32_t)(((2 * kq) ^ (row & 7)) << 4), dqs_c1 = (uint32_t)(((2 * kq + 1) ^ (row & 7)) << 4);
        // The readout of pair u is split around the math of pair u+2: the TMEM load rides with the S / dP loads (its
        // barrier completed long before), the staging stores share the fence of the dS / P~ stores.
        uint32_t dqv[8];
        auto dq_load = [&](int u) {
            mbar_wait(dq_full + u % 3, (u / 3) & 1);   // (also: staging tile u&1 has been stored, see the MMA warp)
            tc_fence_after();
            tmem_ld8(tmem_dq + (u % 3) * kD + lane_addr + kq * 8, dqv);
        };
        auto dq_stage = [&](int u) {
            uint8_t* dst = dqs_row + (u & 1) * (kT * kD * 4);
            *reinterpret_cast<uint4*>(dst + dqs_c0) = make_uint4(dqv[0], dqv[1], dqv[2], dqv[3]);
            *reinterpret_cast<uint4*>(dst + dqs_c1) = make_uint4(dqv[4], dqv[5], dqv[6], dqv[7]);
        };
        auto dq_readout = [&](int u) {   // whole readout (the last two pairs, after the loop)
            dq_load(u);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(dq_empty + u % 3);
            dq_stage(u);
            fence_proxy_async_smem();
            mbar_arrive(dqs_full + (u & 1));
        };

        // per-row scalars of the NEXT tile are fetched one iteration ahead
        float lse_n = CUDART_INF_F, dl_n = 0.f;
        if (row < p.L) { lse_n = lse_bh[row]; dl_n = dl_bh[row]; }

        // The two 64-key halves of a pair are independent down to the accumulation MMAs: start the second group of 8
        // warps about half a period late so that one group's latency-bound steps (barrier waits, TMEM loads, fences)
        // run under the other group's issue-bound math.
#ifndef DETR_BWD_SKEW_NS
#define DETR_BWD_SKEW_NS 400
#endif
        if ((kq >> 1) && T > 1) __nanosleep(DETR_BWD_SKEW_NS);
        uint32_t s[32], dp[32];
        for (int t = 0; t < T; ++t) {
            const int buf = t & 1;
            const int q = t * kT + row;
            // +inf (padding row / fully masked row) -> p = 0.  With dropout P comes out pre-scaled by 1/(1-p) and D is
            // scaled by (1-p) instead:  dS = P/(1-p) o (keep o dP~ - (1-p) D)
            const float nl = drop ? fmaf(-lse_n, 1.4426950408889634f, p.drop_log2_scale) : -lse_n * 1.4426950408889634f;
            const float dlt = drop ? dl_n * p.drop_keep : dl_n;
            {
                const int qn = q + kT;
                lse_n = CUDART_INF_F; dl_n = 0.f;
                if (t + 1 < T && qn < p.L) { lse_n = lse_bh[qn]; dl_n = dl_bh[qn]; }
            }
            const uint32_t row_key = drop ? dropout_row_key(seed, bh, (uint32_t)q) : 0u;
            const uint8_t* arow = (p.amask && q < p.L) ? p.amask + (int64_t)q * p.S + key0 : nullptr;

            BWD_STAMP(0, t);
            mbar_wait(sdp_full + (kq >> 1), t & 1);
            BWD_STAMP(1, t);
            tc_fence_after();
            tmem_ld32(tmem_s + lane_addr + kq * 32, s);
            tmem_ld32(tmem_dp + lane_addr + kq * 32, dp);
            if (t >= 2) dq_load(t - 2);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(sdp_empty + (kq >> 1));                  // this half's score columns may be overwritten by the next pair
            if (t >= 2) mbar_arrive(dq_empty + (t - 2) % 3);
            // dS / P~ buffer `buf` is free: the MMAs of pair t-2 completed before dq_full(t-2), awaited just above
            BWD_STAMP(3, t);

            uint8_t* ds_row = smem + Smem::ds + buf * (kT * kT * 2) + row_off;
            uint8_t* pt_row = smem + Smem::pt + buf * (kT * kT * 2) + row_off;
            // Packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2) halves the floating-point issue slots.  Dropout never
            // touches an fp32 value: a dropped key's dS is -P*D', so both candidates x*(dP - D') and -x*D' are rounded
            // to bf16 pairs and ONE bit-select per pair picks by the keep mask (the same mask clears P~).
            auto quarter_tile = [&](auto masked_c, auto drop_c) {
                constexpr bool MASKED = decltype(masked_c)::value, DROP = decltype(drop_c)::value;
                uint32_t rng = DROP ? dropout_group_state(row_key, (uint32_t)(key0 >> 5)) : 0u;
                const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nl2 = make_float2(nl, nl), ndl2 = make_float2(-dlt, -dlt);
#pragma unroll
                for (int g = 0; g < 4; ++g) {               // 8 keys = one 16-byte chunk of the dS / P~ rows
                    uint32_t t0 = 0, t1 = 0;
                    if (DROP) { t0 = dropout_quad(rng, thr4); t1 = dropout_quad(rng, thr4); }
                    uint32_t pk[4], dk[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = g * 8 + 2 * j;
                        const float2 e = __ffma2_rn(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), sc2, nl2);
                        float2 x = make_float2(ex2(e.x), ex2(e.y));
                        if (MASKED) {
                            const bool m0 = kf[i] != 0 || (arow != nullptr && (key0 + i) < p.S && arow[i] != 0);
                            const bool m1 = kf[i + 1] != 0 || (arow != nullptr && (key0 + i + 1) < p.S && arow[i + 1] != 0);
                            x.x = m0 ? 0.f : x.x; x.y = m1 ? 0.f : x.y;
                        }
                        // dS without the 1/sqrt(d) factor: it is applied once to dK in the epilogue and to dQ in the reduction
                        const float2 a = __fmul2_rn(x, __fadd2_rn(make_float2(__uint_as_float(dp[i]), __uint_as_float(dp[i + 1])), ndl2));
                        pk[j] = pack_bf16x2(x.x, x.y);
                        dk[j] = pack_bf16x2(a.x, a.y);
                        if (DROP) {
                            const float2 bb = __fmul2_rn(x, ndl2);
                            const uint32_t bk = pack_bf16x2(bb.x, bb.y);
                            const uint32_t m = (j & 1) ? dropout_mask_bf16x2<1>(j < 2 ? t0 : t1) : dropout_mask_bf16x2<0>(j < 2 ? t0 : t1);
                            dk[j] = (dk[j] & m) | (bk & ~m);
                            pk[j] &= m;
                        }
                    }
                    const uint32_t off = ((chunk0 + g) ^ (uint32_t)(row & 7)) << 4;
                    *reinterpret_cast<uint4*>(ds_row + off) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
                    *reinterpret_cast<uint4*>(pt_row + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            };
            using TT = std::true_type; using FF = std::false_type;
            if (any) {
                if (drop) quarter_tile(TT{}, TT{}); else quarter_tile(TT{}, FF{});
            } else {
                if (drop) quarter_tile(FF{}, TT{}); else quarter_tile(FF{}, FF{});
            }
            if (t >= 2) dq_stage(t - 2);
            fence_proxy_async_smem();
            mbar_arrive(ds_full + buf);
            if (t >= 2) mbar_arrive(dqs_full + (t & 1));
            BWD_STAMP(4, t);
            BWD_STAMP(5, t);
        }
        BWD_STAMP(6, 0);
        if (T >= 2) dq_readout(T - 2);
        dq_readout(T - 1);
        BWD_STAMP(7, 0);
        // ---- epilogue: dK (key quarters 0,1: 16 columns each) and dV (quarters 2,3) -> bf16 global ----
        // (dq_full of the last pair was committed after every MMA of the stream: the accumulators are complete)
        {
            const int gr = k0 + row;
            const bool is_dk = kq < 2;
            const int c0 = (kq & 1) * 16;
            uint32_t a[16];
            tc_fence_after();
            tmem_ld16((is_dk ? tmem_dk : tmem_dv) + lane_addr + c0, a);
            tmem_ld_wait();
            if (gr < p.S) {
                __nv_bfloat16* dst = is_dk ? p.dk + b * p.dk_sb + (int64_t)gr * p.dk_sl + h * kD + c0
                                           : p.dv + b * p.dv_sb + (int64_t)gr * p.dv_sl + h * kD + c0;
                const float f = is_dk ? p.scale : 1.f;   // dK carries the 1/sqrt(d) of the scores; dV does not
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    uint4 w;
                    w.x = pack_bf16x2(f * __uint_as_float(a[g * 8 + 0]), f * __uint_as_float(a[g * 8 + 1]));
                    w.y = pack_bf16x2(f * __uint_as_float(a[g * 8 + 2]), f * __uint_as_float(a[g * 8 + 3]));
                    w.z = pack_bf16x2(f * __uint_as_float(a[g * 8 + 4]), f * __uint_as_float(a[g * 8 + 5]));
                    w.w = pack_bf16x2(f * __uint_as_float(a[g * 8 + 6]), f * __uint_as_float(a[g * 8 + 7]));
                    reinterpret_cast<uint4*>(dst)[g] = w;
                }
            }
        }
        tc_fence_before();
        BWD_STAMP(6, 1);
    }
    BWD_STAMP(7, 1);
    __syncthreads();
    if (warp == 17) tmem_dealloc(tmem_base, kTmemCols);
#ifdef DETR_BWD_TIMELINE
    if (p.dbg != nullptr && tid == 0 && cta_lin < 1024) {
        long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        p.dbg[2560 + cta_lin * 4 + 2] = gt;
        if (cta_lin == 0) p.dbg[(19 * 16 + 0) * 8 + 1] = clock64();
    }
#endif
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

// dQ[b,q,c] = scale * sum_kt part[kt][b][q][c]   (fixed summation order; 8 channels per thread)
__global__ void attention_dq_reduce_kernel(const float* __restrict__ part, int KT, int64_t part_stride /* B*L*C */,
                                           __nv_bfloat16* __restrict__ dq, int64_t dq_sb, int64_t dq_sl, int B, int L, int C, float scale) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over B*L*C/8
    const int c8 = C >> 3;
    if (idx >= (int64_t)B * L * c8) return;
    const int c = (int)(idx % c8) * 8;
    const int64_t bq = idx / c8;
    const int q = (int)(bq % L), b = (int)(bq / L);
    const float* src = part + ((int64_t)b * L + q) * C + c;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int kt = 0; kt < KT; ++kt) {
        const float4 x = *reinterpret_cast<const float4*>(src + kt * part_stride);
        const float4 y = *reinterpret_cast<const float4*>(src + kt * part_stride + 4);
        acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
        acc[4] += y.x; acc[5] += y.y; acc[6] += y.z; acc[7] += y.w;
    }
    uint4 w;
    w.x = pack_bf16x2(scale * acc[0], scale * acc[1]); w.y = pack_bf16x2(scale * acc[2], scale * acc[3]);
    w.z = pack_bf16x2(scale * acc[4], scale * acc[5]); w.w = pack_bf16x2(scale * acc[6], scale * acc[7]);
    *reinterpret_cast<uint4*>(dq + b * dq_sb + (int64_t)q * dq_sl + c) = w;
}

}  // namespace bwd
}  // namespace detr

using namespace detr;

static long long* g_bwd_dbg = nullptr;
/* debugging aid (not part of the drop-in surface): device buffer of 20*16*8 int64 that receives clock64 stamps of CTA 0 */
extern "C" void detr_attention_bwd_set_debug(long long* buf) { g_bwd_dbg = buf; }

extern "C" int64_t detr_attention_bwd_workspace_floats(int B, int nh, int L, int S) {
    const int64_t kt = (S + bwd::kT - 1) / bwd::kT;
    return kt * B * (int64_t)L * nh * bwd::kD;
}

extern "C" int detr_attention_bwd_bf16(const void* q, int64_t q_sb, int64_t q_sl, const void* k, int64_t k_sb, int64_t k_sl,
                                       const void* v, int64_t v_sb, int64_t v_sl, const void* o, int64_t o_sb, int64_t o_sl,
                                       const void* d_o, int64_t do_sb, int64_t do_sl, const float* lse, float* delta,
                                       float* dq_partial, void* dq, int64_t dq_sb, int64_t dq_sl, void* dk, int64_t dk_sb, int64_t dk_sl,
                                       void* dv, int64_t dv_sb, int64_t dv_sl, const uint8_t* key_padding_mask, int64_t kpm_sb,
                                       const uint8_t* attention_mask, int B, int nh, int L, int S, float dropout_p,
                                       uint64_t seed, const uint64_t* seed_ptr, void* stream) {
    using namespace detr::bwd;
    DETR_CHECK_ARG(B >= 1 && nh >= 1 && L >= 1 && S >= 1, "attention_bwd: bad sizes B=%d nh=%d L=%d S=%d", B, nh, L, S);
    DETR_CHECK_ARG(B <= 65535 && nh <= 65535, "attention_bwd: B and nh must fit the grid");
    DETR_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "attention_bwd: dropout_p must be in [0,1)");
    DETR_CHECK_ARG(dq_partial != nullptr && ((uintptr_t)dq_partial % 128) == 0, "attention_bwd: dq_partial workspace missing or misaligned");
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
    p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.dk_sb = dk_sb; p.dk_sl = dk_sl;
    p.dv = reinterpret_cast<__nv_bfloat16*>(dv); p.dv_sb = dv_sb; p.dv_sl = dv_sl;
    p.dq_part = dq_partial;
    p.lse = lse; p.delta = delta; p.kpm = key_padding_mask; p.kpm_sb = kpm_sb; p.amask = attention_mask;
    p.B = B; p.nh = nh; p.L = L; p.S = S;
    p.scale = 1.f / sqrtf((float)kD);
    p.scale_log2 = 1.4426950408889634f * p.scale;
    p.drop_thresh = (uint32_t)lrintf(dropout_p * 128.f);
    p.drop_keep = (128.f - (float)p.drop_thresh) / 128.f;
    p.drop_log2_scale = -log2f(p.drop_keep);
    p.seed = seed; p.seed_ptr = seed_ptr;
    p.dbg = g_bwd_dbg;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e1 = cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem::total);
        if (e1 != cudaSuccess) { set_error("attention_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e1)); return 2; }
        attr_set = true;
    }
    const int KT = (S + kT - 1) / kT;
    CUtensorMap tdqp;   // partials [KT*B][L][C] fp32, box = 32 channels x 128 queries, SWIZZLE_128B (128-byte rows)
    if (int rc = make_f32_tile_map(&tdqp, dq_partial, C, L, KT * B, kD, kT, "attention_bwd(dQ partials)")) return rc;
    attention_bwd_kernel<<<dim3(KT, nh, B), kThreads, Smem::total, st>>>(tq, tk, tv, tdo, tdqp, p);
    DETR_CHECK_LAUNCH("attention_bwd");
    const int64_t n8 = (int64_t)B * L * (C / 8);
    attention_dq_reduce_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(
        dq_partial, KT, (int64_t)B * L * C, reinterpret_cast<__nv_bfloat16*>(dq), dq_sb, dq_sl, B, L, C, p.scale);
    DETR_CHECK_LAUNCH("attention_dq_reduce");
    return 0;
}
