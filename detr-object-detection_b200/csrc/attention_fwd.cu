// Flash attention forward for DETR's head_dim = 32 (detr/model.py:317-352): TMA-fed tcgen05 with TMEM accumulators.
//
//   O[b,q,h,:] = dropout(softmax(Q K^T / sqrt(32) + mask)) V        LSE[b,h,q] saved for the backward kernel
//
// PERSISTENT: min(#SMs, #items) CTAs; an item is a (batch, head, 128-query tile) and has KT (query tile, key tile)
// pairs; CTA c walks the contiguous pair range [c*N/G, (c+1)*N/G) in (b, h, query tile, key tile) order -- whole items
// plus at most one partial item at each end, so all SMs finish together (448 items on 296 CTA slots used to cost 2 waves
// for 1.51 waves of work).  A partial item leaves (unnormalised O, row max, row sum) in one of two fp32 slots and
// attention_fwd_combine_kernel merges the two parts.
// 20 warps: 0-15 softmax (warp w: TMEM lanes = query rows 32*(w%4).., key columns 32*(w/4)..+31 of each 128x128 score tile,
// and 8 of the 32 output columns), 16 = TMA producer, 17 = tcgen05.mma issuer, 18-19 idle (fifth warpgroup for setmaxnreg).
// Per pair j (global counter across items) the only barrier a softmax warp waits on is s_full[j&1]:
//   MMA warp:  wait p_full[j&1] -> O_tile[j&1] = P_j V_j -> S[(j+2)&1] = Q K_{j+2}^T -> commit s_full[(j+2)&1]
//   so "S_{j+2} is ready" also means "P V of pair j is complete": the P buffer and the O_tile columns of pair j are free /
//   readable, and S_j had been copied to registers before p_full(j) was signalled.  The output accumulates in registers
//   two pairs late:  acc = (acc + O_tile(j-2) * alpha(j-1)) * alpha(j).
// TMEM: S x2 [0,256) | O_tile x2 [256,320).  Shared memory: Q x2 16 KB | K/V ring 4 x 16 KB | P x2 64 KB | row-max exchange | deferred-finish stash 24 KB.
//
// Masking follows the reference: key_padding_mask / attention_mask entries get a huge FINITE negative score
// (detr/model.py:326-334 uses finfo.min, so a fully masked row is uniform, not NaN); keys beyond S (tile
// padding) get -inf and never contribute.
#include <math_constants.h>

#include "common.cuh"
#include "tc.cuh"

namespace detr {
using namespace tc;

constexpr int kBM = 128;          // queries per item
constexpr int kBN = 128;          // keys per pair
constexpr int kD = 32;            // head dim
constexpr int kStages = 4;
constexpr int kSoftmaxWarps = 16;
constexpr int kSoftmaxThreads = kSoftmaxWarps * 32;
constexpr int kFwdThreads = 640;
constexpr uint32_t kTileBytes = kBN * kD * 2;  // 8 KB: one Q, K or V tile
constexpr uint32_t kTmemCols = 512;
// Masked score: huge, finite, and a POWER OF TWO (-2^126) so that score*scale is exact and the fused
// multiply-add s*scale - m*scale cancels to exactly 0 when a whole row is masked (-> uniform probabilities).
constexpr float kMaskedScore = -8.507059173023462e37f;

// EXPERIMENT, off by default: the 8-key groups selected by kPolyGroups (bit g = group g of a thread's 32 keys) take their
// exponentials on the FMA pipe instead of the exp unit (MUFU, 4 lanes per scheduler):
// 2^x = 2^n * p(f), n = round(x), f = x - n in [-0.5, 0.5], p = degree-3 minimax polynomial (relative error 7.5e-5, far
// below the bf16 rounding of P), n added into the exponent field, packed fp32x2 arithmetic throughout.
// Measured on B200 (config 2 encoder shape / DC5): groups 0xA (half of the keys) 50.2 -> 52.2 us / 126 -> 138 us: the
// softmax warps are bound by issue slots (4 warps per scheduler, 54 % issue-active, XU 35 %: profiles/r01_attention_ncu_full_v4.md),
// so trading 1 MUFU for ~5 FMA-pipe instructions loses.  Kept for head dims / shapes where the balance differs.
#ifndef DETR_FWD_POLY_GROUPS
#define DETR_FWD_POLY_GROUPS 0x0
#endif
constexpr uint32_t kPolyGroups = DETR_FWD_POLY_GROUPS;
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
    const float kMagic = 12582912.f;   // 1.5 * 2^23: x + kMagic holds round(x) in its low mantissa bits
    x.x = fmaxf(x.x, -125.f); x.y = fmaxf(x.y, -125.f);   // masked / out-of-range keys: 2^-125 instead of 0
    const float2 t = __fadd2_rn(x, make_float2(kMagic, kMagic));
    const float2 n = __fadd2_rn(t, make_float2(-kMagic, -kMagic));
    const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);
    float2 r = __ffma2_rn(f, make_float2(0.05517143756151199f, 0.05517143756151199f), make_float2(0.24261081218719482f, 0.24261081218719482f));
    r = __ffma2_rn(r, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
    r = __ffma2_rn(r, f, make_float2(0.9999281167984009f, 0.9999281167984009f));
    r.x = __int_as_float(__float_as_int(r.x) + (__float_as_int(t.x) << 23));
    r.y = __int_as_float(__float_as_int(r.y) + (__float_as_int(t.y) << 23));
    return r;
}

struct AttnFwdParams {
    __nv_bfloat16* O; int64_t o_sb, o_sl;  // (B, L, nh*32): element strides of batch and row
    __nv_bfloat16* O_lo;                   // optional, same strides: bf16(O_fp32 - bf16(O_fp32)), the rounding residual of O.  The backward
                                           // pass computes delta = rowsum(dO * (O + O_lo)): with the rounded O alone the error of delta is common to all
                                           // keys of a row and does not average out in dQ / dK when the attention is nearly uniform
    float* lse;                            // (B, nh, L)
    float* part;                           // [items][2 slots][128 rows][36]: unnormalised O (32), row max, row sum of split items
    const uint8_t* kpm; int64_t kpm_sb;    // key padding mask (B, S) bytes, may be null
    const uint8_t* amask;                  // attention mask (L, S) bytes, may be null
    int B, nh, L, S;
    float scale_log2;                      // log2(e) / sqrt(32)
    uint32_t drop_thresh;                  // 0 = no dropout; a key is dropped if its 15 random bits < thresh
    float drop_scale;                      // 32768 / (32768 - thresh)
    float drop_log2_scale;                 // log2(drop_scale): folded into the exponent
    uint64_t seed; const uint64_t* seed_ptr; // effective seed = seed + *seed_ptr (device side: CUDA-graph replays get fresh masks)
    long long* dbg;                        // optional clock64 timeline of CTA 0 (DETR_FWD_TIMELINE builds), NULL in production
};
constexpr int kPartRow = 36;   // 32 output columns, row max, row sum, padding to a 16-byte multiple

struct FwdSmem {
    static constexpr uint32_t q = 0;                                     // 2 x Q tile: the next item's Q arrives while the current one is in use
    static constexpr uint32_t kv = q + 2 * kTileBytes;                   // kStages x (K tile, V tile)
    static constexpr uint32_t p = 81920;                                 // 2 x (128 x 128 bf16, two 64-key blocks, SWIZZLE_128B)
    static constexpr uint32_t xch = p + 2 * kBM * kBN * 2;               // float[2 parity + 1 (row sums)][4 key quarters][128 rows]
    static constexpr uint32_t bars = xch + 3 * 4 * kBM * 4;
    static constexpr uint32_t stash = bars + 256;                        // float4[512 threads][3]: accumulator, alpha, row max of a segment whose finish is deferred
    static constexpr uint32_t total = stash + kSoftmaxThreads * 48 + 1024;   // + alignment slack
};
static_assert(FwdSmem::kv + kStages * 2 * kTileBytes <= FwdSmem::p && FwdSmem::p % 1024 == 0, "smem layout");

// persistent schedule (same scheme as the backward kernel): item = ((b * nh + h) * QT + qt), T = KT pairs per item
struct FwdSched {
    int T, QT, nh, NT, item0, t0;
    __device__ __forceinline__ void init(const AttnFwdParams& p, int c, int G) {
        T = (p.S + kBN - 1) / kBN; QT = (p.L + kBM - 1) / kBM; nh = p.nh;
        const long long total = (long long)QT * p.nh * p.B * T;
        const long long n0 = total * c / G, n1 = total * (c + 1) / G;
        NT = (int)(n1 - n0); item0 = (int)(n0 / T); t0 = (int)(n0 - (long long)item0 * T);
    }
    __device__ __forceinline__ void split(int item, int& b, int& h, int& qt) const { qt = item % QT; const int bh = item / QT; h = bh % nh; b = bh / nh; }
};
struct FwdCursor {
    int item, t;
    __device__ __forceinline__ explicit FwdCursor(const FwdSched& sc) : item(sc.item0), t(sc.t0) {}
    __device__ __forceinline__ void next(const FwdSched& sc) { if (++t == sc.T) { t = 0; ++item; } }
};

#ifndef DETR_FWD_TIMELINE
#define FWD_STAMP(ev, j) do {} while (0)
#else
#define FWD_STAMP(ev, j) do { if (p.dbg != nullptr && lane == 0 && blockIdx.x == 0 && (j) < 32) \
    p.dbg[(warp * 32 + (j)) * 8 + (ev)] = clock64(); } while (0)
#endif

// bf16 of (o[e] * inv - what the packed bf16 words w hold): the part of the output lost to rounding
__device__ __forceinline__ uint4 residual_bf16x8(const float (&o)[8], float inv, const uint4& w) {
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
    uint32_t r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        r[j] = pack_bf16x2(o[2 * j] * inv - __uint_as_float(ww[j] << 16), o[2 * j + 1] * inv - __uint_as_float(ww[j] & 0xffff0000u));
    return make_uint4(r[0], r[1], r[2], r[3]);
}

__global__ void __launch_bounds__(kFwdThreads, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const AttnFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzled tiles need 1024-byte alignment
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    FwdSched sc;
    sc.init(p, blockIdx.x, gridDim.x);
    const int NT = sc.NT;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::bars);
    uint64_t* q_full = bars + 0;                // [2]
    uint64_t* q_empty = bars + 2;               // [2]
    uint64_t* kv_full = bars + 4;               // [kStages]
    uint64_t* kv_empty = kv_full + kStages;     // [kStages]
    uint64_t* s_full = kv_empty + kStages;      // [2]
    uint64_t* p_full = s_full + 2;              // [2]
    uint64_t* o_done = s_full + 4;              // all P V of a segment complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 5);

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(q_full + s, 1); mbar_init(q_empty + s, 1); }
        mbar_init(o_done, 1);
        for (int s = 0; s < kStages; ++s) { mbar_init(kv_full + s, 1); mbar_init(kv_empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(s_full + s, 1); mbar_init(p_full + s, kSoftmaxThreads); }
        fence_barrier_init();
    }
    if (warp == 17) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 256;
    pdl_wait();      // small calls are launched programmatically behind the projection GEMM: the prologue above overlaps its tail
    pdl_trigger();   // the combine launch may be scheduled as SMs free up (it waits for this grid before reading the partial slots)

    if (warp >= kSoftmaxWarps) {
        reg_dealloc<64>();
        if (warp == 16 && lane == 0) {
            // ================= TMA producer =================
            tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
            int seg = -1, b = 0, h = 0, qt = 0;
            FwdCursor cur(sc);
            for (int j = 0; j < NT; ++j, cur.next(sc)) {
                if (j == 0 || cur.t == 0) {   // new segment: its Q tile goes to buffer seg & 1, free once the last score MMAs of segment seg-2 have
                    ++seg;                    // read it -- long ago: with a single buffer the producer stalled here until the END of the previous
                    sc.split(cur.item, b, h, qt);   // segment and the K/V ring drained behind it (~5 000 cycles per item boundary)
                    if (seg >= 2) mbar_wait_sleep(q_empty + (seg & 1), ((seg >> 1) - 1) & 1);
                    mbar_expect_tx(q_full + (seg & 1), kTileBytes);
                    tma_load_3d(smem + FwdSmem::q + (seg & 1) * kTileBytes, &tm_q, q_full + (seg & 1), h * kD, qt * kBM, b);
                }
                const int st = j % kStages;
                if (j >= kStages) mbar_wait_sleep(kv_empty + st, ((j / kStages) - 1) & 1);
                mbar_expect_tx(kv_full + st, 2 * kTileBytes);
                uint8_t* dst = smem + FwdSmem::kv + st * 2 * kTileBytes;
                tma_load_3d(dst, &tm_k, kv_full + st, h * kD, cur.t * kBN, b);
                tma_load_3d(dst + kTileBytes, &tm_v, kv_full + st, h * kD, cur.t * kBN, b);
            }
        } else if (warp == 17 && elect_one()) {
            // ================= MMA issuer =================
            constexpr uint32_t idesc_s = make_idesc_bf16(kBM, kBN, false, false);
            constexpr uint32_t idesc_o = make_idesc_bf16(kBM, kD, false, true);
            // descriptor words (tc.cuh): high word per layout, low word = (address >> 4) + constant
            constexpr uint32_t hi64 = desc_hi(512, SWZ_64B);      // Q/K/V tiles: 64-byte rows, 8-row groups 512 B apart
            constexpr uint32_t hi128 = desc_hi(1024, SWZ_128B);   // P tile: 128-byte rows, 8-row groups 1024 B apart
            const uint32_t q_lo0 = smem_u32(smem + FwdSmem::q) >> 4;
            // `seg_of_scores`: segment of the last score MMA issued; scores of a new segment wait for its Q tile
            int seg_scores = -1;
            FwdCursor sc_cur(sc);   // cursor of the NEXT pair whose scores will be issued
            int js = 0;             // its local index
            auto scores = [&]() {
                const int j = js;
                const bool first = j == 0 || sc_cur.t == 0, last = j == NT - 1 || sc_cur.t == sc.T - 1;
                if (first) { ++seg_scores; mbar_wait_sleep(q_full + (seg_scores & 1), (seg_scores >> 1) & 1); }
                const uint32_t q_lo = q_lo0 + (uint32_t)(seg_scores & 1) * (kTileBytes >> 4);
                mbar_wait_sleep(kv_full + (j % kStages), (j / kStages) & 1);
                tc_fence_after();
                const uint32_t k_lo = smem_u32(smem + FwdSmem::kv + (j % kStages) * 2 * kTileBytes) >> 4;
#pragma unroll
                for (int ks = 0; ks < kD / 16; ++ks)  // K-major, 64-byte rows, SWIZZLE_64B: 32 B per 16-channel step
                    umma_bf16_lh(tmem_s + (j & 1) * kBN, q_lo + desc_lo(ks * 32, 16), hi64, k_lo + desc_lo(ks * 32, 16), hi64, idesc_s, ks > 0);
                umma_commit(s_full + (j & 1));
                if (last) umma_commit(q_empty + (seg_scores & 1));   // no later MMA of this segment reads the Q tile
                ++js; sc_cur.next(sc);
            };
            if (NT > 0) scores();
            if (NT > 1) scores();
            FwdCursor cur(sc);
            for (int j = 0; j < NT; ++j, cur.next(sc)) {
                const bool last = j == NT - 1 || cur.t == sc.T - 1;
                FWD_STAMP(0, j);
                mbar_wait_sleep(p_full + (j & 1), (j >> 1) & 1);     // P_j is in shared memory (and S_j, O_tile(j-2) are in registers)
                FWD_STAMP(1, j);
                tc_fence_after();
                const uint32_t v_lo = (smem_u32(smem + FwdSmem::kv + (j % kStages) * 2 * kTileBytes) + kTileBytes) >> 4;
                const uint32_t p_lo = smem_u32(smem + FwdSmem::p + (j & 1) * (kBM * kBN * 2)) >> 4;
#pragma unroll
                for (int ks = 0; ks < kBN / 16; ++ks) {
                    // A = P: K-major SWIZZLE_128B, 64-key blocks of 16 KB, 32 B per 16-key step inside a block
                    // B = V: MN-major (d contiguous, 64-byte rows), SWIZZLE_64B, 16 keys = 1024 B per step
                    umma_bf16_lh(tmem_o + (j & 1) * kD, p_lo + desc_lo((ks >> 2) * 16384 + (ks & 3) * 32, 16), hi128,
                                 v_lo + desc_lo(ks * 1024, 512), hi64, idesc_o, ks > 0);
                }
                umma_commit(kv_empty + (j % kStages));
                if (last) umma_commit(o_done);
                if (j + 2 < NT) scores();                          // S_{j+2}: its commit also covers P V of pair j
                FWD_STAMP(2, j);
            }
        }
    } else {
        // ================= softmax warps =================
        reg_alloc<104>();
        const int lq = warp & 3, kq = warp >> 2;      // TMEM lane quarter, 32-key quarter of the pair
        const int row = lq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
        const bool drop = p.drop_thresh != 0;
        const uint64_t seed = drop ? p.seed + (p.seed_ptr ? *p.seed_ptr : 0ull) : 0ull;
        const uint32_t thr2 = p.drop_thresh * 0x00010001u;
        const float sc_l2 = p.scale_log2;
        // this thread's row inside a [query][64-key block] SWIZZLE_128B tile; its 4 chunks are (kq&1)*4 + g
        const uint32_t row_off = (uint32_t)((kq >> 1) * 16384 + (row >> 3) * 1024 + (row & 7) * 128);
        const uint32_t chunk0 = (uint32_t)((kq & 1) * 4);
        float* xch = reinterpret_cast<float*>(smem + FwdSmem::xch);

        int seg = -1, seg_t0 = 0, b = 0, h = 0, q = 0;
        uint32_t bh = 0, row_key = 0;
        float m_run = -CUDART_INF_F, l_run = 0.f, alpha_prev = 1.f;
        float acc[8];
        uint8_t pad_next = 0;
        FwdCursor cur(sc);

        // normalised output + LSE of a whole item, or the (unnormalised O, row max, row sum) partial slot of a split one
        auto store_out = [&](const float (&o)[8], float l_tot, float m, int b_, int h_, int q_, int item, bool whole, int slot) {
            if (q_ >= p.L) return;
            if (whole) {
                const float inv = 1.f / l_tot;
                uint4 w;
                w.x = pack_bf16x2(o[0] * inv, o[1] * inv); w.y = pack_bf16x2(o[2] * inv, o[3] * inv);
                w.z = pack_bf16x2(o[4] * inv, o[5] * inv); w.w = pack_bf16x2(o[6] * inv, o[7] * inv);
                const int64_t off = b_ * p.o_sb + (int64_t)q_ * p.o_sl + h_ * kD + kq * 8;
                *reinterpret_cast<uint4*>(p.O + off) = w;
                if (p.O_lo != nullptr) *reinterpret_cast<uint4*>(p.O_lo + off) = residual_bf16x8(o, inv, w);
                // natural-log LSE of the scaled scores: m/sqrt(d) + ln(l).  A row whose keys are all masked is flagged
                // with +inf: the backward kernel then skips it.
                if (kq == 0)
                    p.lse[((int64_t)b_ * p.nh + h_) * p.L + q_] =
                        m == kMaskedScore ? CUDART_INF_F : (m * sc_l2 + log2f(l_tot)) * 0.6931471805599453f;
            } else {
                float* dst = p.part + (((int64_t)item * 2 + slot) * kBM + row) * kPartRow;
                *reinterpret_cast<float4*>(dst + kq * 8) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(dst + kq * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
                if (kq == 0) { dst[32] = m; dst[33] = l_tot; }
            }
        };
        // total row sum = sum over the 4 key quarters (same running max, same alpha history), from the third exchange buffer
        auto row_sum_total = [&]() {
            const float* xr = xch + 2 * 512 + row;
            float l_tot = (xr[0] + xr[128]) + (xr[256] + xr[384]);
            if (drop) l_tot *= 1.f / p.drop_scale;              // the exponentials were pre-scaled by 1/(1-p)
            return l_tot;
        };
        // finish a segment NOW: wait for its last P V, read the two outstanding O tiles, exchange the row sums, store.  Used
        // where the finish cannot be deferred (see `pend` below): it costs ~5 000 cycles during which these warps start nothing.
        auto finish = [&](int j_last, int item, bool whole, int slot, int n_tiles) {
            mbar_wait(o_done, seg & 1);
            tc_fence_after();
            uint32_t o0[8], o1[8];
            if (n_tiles >= 2) tmem_ld8(tmem_o + ((j_last - 1) & 1) * kD + lane_addr + kq * 8, o0);
            tmem_ld8(tmem_o + (j_last & 1) * kD + lane_addr + kq * 8, o1);
            // (third exchange buffer: a fast warp may already be writing the next pair's row max into either parity buffer)
            xch[2 * 512 + kq * 128 + row] = l_run;
            named_bar_sync(1 + lq, 128);
            const float l_tot = row_sum_total();
            tmem_ld_wait();
            tc_fence_before();
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float v = acc[e];
                if (n_tiles >= 2) v += __uint_as_float(o0[e]) * alpha_prev;   // O_tile(last-1) is relative to the max before the last pair
                o[e] = v + __uint_as_float(o1[e]);
            }
            store_out(o, l_tot, m_run, b, h, q, item, whole, slot);
        };
        // DEFERRED finish.  A segment that ends at pair j with at least three more pairs of the NEXT item behind it in this CTA
        // does not wait: its accumulator, last alpha and row max go to the stash, its row sum to the third exchange buffer, and
        // the two outstanding O tiles are picked up by the loads the next two pairs issue anyway (s_full(j+1) implies
        // P V(j-1) complete, s_full(j+2) implies P V(j) complete; the new item has no O tile of its own to read there).
        // pend: 0 = nothing, 2 = O_tile(last-1) comes with the next pair, 1 = O_tile(last) comes with the next pair.
        // pend_info = item << 3 | whole << 2 | slot << 1 | (n_tiles >= 2).
        int pend = 0, pend_info = 0;
        float4* stash = reinterpret_cast<float4*>(smem + FwdSmem::stash) + tid * 3;

        uint32_t s[32];
        for (int j = 0; j < NT; ++j, cur.next(sc)) {
            const int t = cur.t;
            if (j == 0 || t == 0) {
                ++seg; seg_t0 = t;
                int qt;
                sc.split(cur.item, b, h, qt);
                q = qt * kBM + row;
                bh = (uint32_t)(b * p.nh + h);
                row_key = drop ? dropout_row_key(seed, bh, (uint32_t)q) : 0u;
                m_run = -CUDART_INF_F; l_run = 0.f; alpha_prev = 1.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = 0.f;
            }
            const int jt = j - (t - seg_t0);                       // local index of the segment's first pair
            const int key0 = t * kBN + kq * 32;
            // key flags of this thread's 32 columns: bit i of `oob` = beyond S (-inf), of `pad` = key_padding_mask (finite).
            // The mask byte of the NEXT pair is fetched one iteration ahead (its L2 latency would sit on the critical path).
            const int n_valid = p.S - key0;
            const uint32_t oob = n_valid >= 32 ? 0u : (n_valid <= 0 ? 0xffffffffu : (0xffffffffu << n_valid));
            if (p.kpm != nullptr && (j == 0 || t == 0)) pad_next = (key0 + lane < p.S) ? p.kpm[b * p.kpm_sb + key0 + lane] : (uint8_t)0;
            const uint32_t pad = p.kpm ? __ballot_sync(FULL_MASK, pad_next != 0) : 0u;
            if (p.kpm != nullptr && t + 1 < sc.T) {
                const int kn = key0 + kBN + lane;
                pad_next = kn < p.S ? p.kpm[b * p.kpm_sb + kn] : (uint8_t)0;
            }
            const uint8_t* arow = (p.amask && q < p.L) ? p.amask + (int64_t)q * p.S + key0 : nullptr;
            const bool any = (oob | pad) != 0u || p.amask != nullptr;

            FWD_STAMP(0, j);
            mbar_wait(s_full + (j & 1), (j >> 1) & 1);            // S_j ready; P V of pair j-2 complete
            FWD_STAMP(1, j);
            tc_fence_after();
            tmem_ld32(tmem_s + (j & 1) * kBN + lane_addr + kq * 32, s);
            uint32_t o_old[8];
            const bool have_old = j - 2 >= jt;                     // O_tile(j-2) belongs to this segment
            if (have_old || pend != 0) tmem_ld8(tmem_o + (j & 1) * kD + lane_addr + kq * 8, o_old);
            tmem_ld_wait();
            tc_fence_before();
            FWD_STAMP(2, j);

            if (any) {
                // bit i of `fin`: finite mask (key padding / attention mask), of `oob`: beyond S.  Two bit tests + selects per
                // score (the per-element short-circuit form cost ~1 400 cycles on the last key tile of every item: timeline)
                uint32_t fin = pad;
                if (arow != nullptr) {
#pragma unroll 4
                    for (int i = 0; i < 32; ++i)
                        if (i < n_valid && arow[i] != 0) fin |= 1u << i;
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float v = __uint_as_float(s[i]);
                    v = (fin & (1u << i)) ? kMaskedScore : v;
                    v = (oob & (1u << i)) ? -CUDART_INF_F : v;
                    s[i] = __float_as_uint(v);
                }
            }
            float mx = fmaxf(__uint_as_float(s[0]), __uint_as_float(s[1]));
#pragma unroll
            for (int i = 2; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(s[i]), __uint_as_float(s[i + 1])));
            // exchange with the three warps that hold the other columns of this row
            xch[(j & 1) * 512 + kq * 128 + row] = mx;
#ifndef DETR_FWD_EXP_NOXCH
            named_bar_sync(1 + lq, 128);
#endif
            {
                const float* xr = xch + (j & 1) * 512 + row;
                mx = fmaxf(fmaxf(xr[0], xr[128]), fmaxf(xr[256], xr[384]));
            }
            FWD_STAMP(3, j);
            const float m_new = fmaxf(m_run, mx);                 // finite: every pair has at least one in-range key
            const float alpha = ex2((m_run - m_new) * sc_l2);     // first pair: exp2(-inf) = 0
            const float bias = drop ? fmaf(-m_new, sc_l2, p.drop_log2_scale) : -m_new * sc_l2;   // kept entries come out pre-scaled by 1/(1-p)
            uint8_t* p_row = smem + FwdSmem::p + (j & 1) * (kBM * kBN * 2) + row_off;
            uint32_t rng = drop ? dropout_group_state(row_key, (uint32_t)(key0 >> 5)) : 0u;
            const float2 sc2 = make_float2(sc_l2, sc_l2), bias2 = make_float2(bias, bias);
            float2 rs = make_float2(0.f, 0.f);
#pragma unroll
            for (int g = 0; g < 4; ++g) {                         // 8 values = one 16-byte chunk of the P row
                uint32_t w[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int i = g * 8 + 2 * jj;
                    const float2 e = __ffma2_rn(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), sc2, bias2);
                    const float2 x = ((kPolyGroups >> g) & 1) ? exp2_poly2(e) : make_float2(ex2(e.x), ex2(e.y));
                    rs = __fadd2_rn(rs, x);
                    w[jj] = pack_bf16x2(x.x, x.y);
                }
                if (drop) {   // the row sum is of the un-dropped probabilities; dropped entries are cleared in the packed bf16 words
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) w[jj] &= dropout_mask_bf16x2(dropout_pair(rng, thr2));
                }
                *reinterpret_cast<uint4*>(p_row + (((chunk0 + g) ^ (uint32_t)(row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            FWD_STAMP(4, j);
            fence_proxy_async_smem();
            mbar_arrive(p_full + (j & 1));
            FWD_STAMP(5, j);
            if (pend != 0) {   // the O tile just read belongs to the previous item (have_old is false in an item's first two pairs)
                float4 a0 = stash[0], a1 = stash[1];
                const float4 a2 = stash[2];   // (last alpha, row max, -, -)
                if (pend == 2) {
                    if (pend_info & 1) {
                        a0.x = fmaf(__uint_as_float(o_old[0]), a2.x, a0.x); a0.y = fmaf(__uint_as_float(o_old[1]), a2.x, a0.y);
                        a0.z = fmaf(__uint_as_float(o_old[2]), a2.x, a0.z); a0.w = fmaf(__uint_as_float(o_old[3]), a2.x, a0.w);
                        a1.x = fmaf(__uint_as_float(o_old[4]), a2.x, a1.x); a1.y = fmaf(__uint_as_float(o_old[5]), a2.x, a1.y);
                        a1.z = fmaf(__uint_as_float(o_old[6]), a2.x, a1.z); a1.w = fmaf(__uint_as_float(o_old[7]), a2.x, a1.w);
                        stash[0] = a0; stash[1] = a1;
                    }
                    pend = 1;
                } else {
                    const float o[8] = {a0.x + __uint_as_float(o_old[0]), a0.y + __uint_as_float(o_old[1]), a0.z + __uint_as_float(o_old[2]),
                                        a0.w + __uint_as_float(o_old[3]), a1.x + __uint_as_float(o_old[4]), a1.y + __uint_as_float(o_old[5]),
                                        a1.z + __uint_as_float(o_old[6]), a1.w + __uint_as_float(o_old[7])};
                    int pb, ph, pqt;
                    sc.split(pend_info >> 3, pb, ph, pqt);
                    store_out(o, row_sum_total(), a2.y, pb, ph, pqt * kBM + row, pend_info >> 3, (pend_info >> 2) & 1, (pend_info >> 1) & 1);
                    pend = 0;
                }
            }
            // output accumulation, two pairs late: acc = (acc + O_tile(j-2) * alpha(j-1)) * alpha(j)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float v = acc[e];
                if (have_old) v = fmaf(__uint_as_float(o_old[e]), alpha_prev, v);
                acc[e] = v * alpha;
            }
            l_run = l_run * alpha + (rs.x + rs.y);                // (pre-scaled by 1/(1-p) when dropping; undone in `finish`)
            m_run = m_new;
            alpha_prev = alpha;
            if (j == NT - 1 || t == sc.T - 1) {
                const int n_tiles = j - jt + 1;
                const bool whole = seg_t0 == 0 && t == sc.T - 1;
                const int slot = seg_t0 == 0 ? 0 : 1;
#ifndef DETR_FWD_NO_DEFER
                // pairs j+1, j+2 exist in this CTA and belong to the next item, and that item cannot end before pair j+3: its own
                // finish (which reuses the third exchange buffer) is separated from the deferred read by a named barrier
                if (sc.T >= 3 && j + 3 <= NT - 1) {
                    stash[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    stash[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
                    stash[2] = make_float4(alpha_prev, m_run, 0.f, 0.f);
                    xch[2 * 512 + kq * 128 + row] = l_run;   // read two pairs (two named barriers) later
                    pend = 2;
                    pend_info = (cur.item << 3) | (whole ? 4 : 0) | (slot << 1) | (n_tiles >= 2 ? 1 : 0);
                } else
#endif
                    finish(j, cur.item, whole, slot, n_tiles);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 17) tmem_dealloc(tmem_base, kTmemCols);
}

// Merge the two parts of the items that were split between two persistent CTAs:
//   m = max(m0, m1), l = l0 2^{(m0-m) s} + l1 2^{(m1-m) s}, O = (O0 2^{(m0-m) s} + O1 2^{(m1-m) s}) / l
// CTA c handles the boundary between persistent CTAs c and c+1 (nothing to do when it falls on an item edge).
__global__ void __launch_bounds__(512) attention_fwd_combine_kernel(const AttnFwdParams p, int G) {
    pdl_wait();
    const int T = (p.S + kBN - 1) / kBN, QT = (p.L + kBM - 1) / kBM;
    const long long total = (long long)QT * p.nh * p.B * T;
    const long long n = total * (blockIdx.x + 1) / G;
    if (n % T == 0) return;
    const int item = (int)(n / T);
    const int qt = item % QT, bh = item / QT, h = bh % p.nh, b = bh / p.nh;
    const int row = threadIdx.x >> 2, g = threadIdx.x & 3, q = qt * kBM + row;   // 4 threads per row, 8 output columns each
    if (q >= p.L) return;
    const float* s0 = p.part + (((int64_t)item * 2 + 0) * kBM + row) * kPartRow;
    const float* s1 = s0 + (int64_t)kBM * kPartRow;
    const float m0 = s0[32], l0 = s0[33], m1 = s1[32], l1 = s1[33];
    const float m = fmaxf(m0, m1);
    const float a0 = ex2((m0 - m) * p.scale_log2), a1 = ex2((m1 - m) * p.scale_log2);
    const float l = l0 * a0 + l1 * a1;
    const float inv = 1.f / l;
    const float4 x0 = *reinterpret_cast<const float4*>(s0 + g * 8), x1 = *reinterpret_cast<const float4*>(s0 + g * 8 + 4);
    const float4 y0 = *reinterpret_cast<const float4*>(s1 + g * 8), y1 = *reinterpret_cast<const float4*>(s1 + g * 8 + 4);
    uint4 w;
    w.x = pack_bf16x2((x0.x * a0 + y0.x * a1) * inv, (x0.y * a0 + y0.y * a1) * inv);
    w.y = pack_bf16x2((x0.z * a0 + y0.z * a1) * inv, (x0.w * a0 + y0.w * a1) * inv);
    w.z = pack_bf16x2((x1.x * a0 + y1.x * a1) * inv, (x1.y * a0 + y1.y * a1) * inv);
    w.w = pack_bf16x2((x1.z * a0 + y1.z * a1) * inv, (x1.w * a0 + y1.w * a1) * inv);
    const int64_t off = b * p.o_sb + (int64_t)q * p.o_sl + h * kD + g * 8;
    *reinterpret_cast<uint4*>(p.O + off) = w;
    if (p.O_lo != nullptr) {
        const float o[8] = {x0.x * a0 + y0.x * a1, x0.y * a0 + y0.y * a1, x0.z * a0 + y0.z * a1, x0.w * a0 + y0.w * a1,
                            x1.x * a0 + y1.x * a1, x1.y * a0 + y1.y * a1, x1.z * a0 + y1.z * a1, x1.w * a0 + y1.w * a1};
        *reinterpret_cast<uint4*>(p.O_lo + off) = residual_bf16x8(o, inv, w);
    }
    if (g == 0)
        p.lse[((int64_t)b * p.nh + h) * p.L + q] = m == kMaskedScore ? CUDART_INF_F : (m * p.scale_log2 + log2f(l)) * 0.6931471805599453f;
}

// ---- host side -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// (C inner, rows, batch) bf16 tensor; box = 32 channels x box_rows rows, SWIZZLE_64B, zero fill out of bounds
int make_head_tile_map(CUtensorMap* out, const void* base, int C, int rows, int B, int64_t row_stride_el, int64_t batch_stride_el,
                       int box_rows, const char* who) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("%s: cuTensorMapEncodeTiled entry point not available", who); return 2; }
    if (((uintptr_t)base % 16) || (row_stride_el % 8) || (batch_stride_el % 8)) {
        set_error("%s: tensor must be 16-byte aligned with row/batch strides that are multiples of 8 elements", who);
        return 1;
    }
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)row_stride_el * 2, (cuuint64_t)batch_stride_el * 2};
    cuuint32_t box[3] = {(cuuint32_t)kD, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r); return 2; }
    return 0;
}

// contiguous fp32 (slabs, rows, C) tensor; box = box_c channels (128 bytes) x box_rows rows, SWIZZLE_128B; used for TMA stores
int make_f32_tile_map(CUtensorMap* out, const void* base, int C, int rows, int slabs, int box_c, int box_rows, const char* who) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("%s: cuTensorMapEncodeTiled entry point not available", who); return 2; }
    if (((uintptr_t)base % 128) || box_c * 4 != 128 || (C % 4)) { set_error("%s: needs a 128-byte aligned base and 128-byte box rows", who); return 1; }
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)slabs};
    cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)rows * C * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r); return 2; }
    return 0;
}

}  // namespace detr

using namespace detr;

static long long* g_fwd_dbg = nullptr;
/* debugging aid (not part of the drop-in surface): device buffer of 20*32*8 int64 that receives clock64 stamps of CTA 0 */
extern "C" void detr_attention_fwd_set_debug(long long* buf) { g_fwd_dbg = buf; }

extern "C" int64_t detr_attention_fwd_workspace_floats(int B, int nh, int L, int S) {
    (void)S;
    const int64_t items = (int64_t)((L + kBM - 1) / kBM) * nh * B;
    return items * 2 * kBM * kPartRow;
}

extern "C" int detr_attention_fwd_bf16(const void* q, int64_t q_sb, int64_t q_sl, const void* k, int64_t k_sb, int64_t k_sl,
                                       const void* v, int64_t v_sb, int64_t v_sl, void* o, int64_t o_sb, int64_t o_sl, void* o_lo,
                                       float* lse, float* workspace, const uint8_t* key_padding_mask, int64_t kpm_sb,
                                       const uint8_t* attention_mask, int B, int nh, int L, int S, float dropout_p,
                                       uint64_t seed, const uint64_t* seed_ptr, void* stream) {
    DETR_CHECK_ARG(B >= 1 && nh >= 1 && L >= 1 && S >= 1, "attention_fwd: bad sizes B=%d nh=%d L=%d S=%d", B, nh, L, S);
    DETR_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "attention_fwd: dropout_p must be in [0,1)");
    DETR_CHECK_ARG(((uintptr_t)o % 16) == 0 && (o_sb % 8) == 0 && (o_sl % 8) == 0, "attention_fwd: O must be 16-byte aligned rows");
    DETR_CHECK_ARG(workspace != nullptr && ((uintptr_t)workspace % 16) == 0, "attention_fwd: workspace missing or misaligned");
    const int C = nh * kD;
    CUtensorMap tq, tk, tv;
    if (int rc = make_head_tile_map(&tq, q, C, L, B, q_sl, q_sb, kBM, "attention_fwd(Q)")) return rc;
    if (int rc = make_head_tile_map(&tk, k, C, S, B, k_sl, k_sb, kBN, "attention_fwd(K)")) return rc;
    if (int rc = make_head_tile_map(&tv, v, C, S, B, v_sl, v_sb, kBN, "attention_fwd(V)")) return rc;
    AttnFwdParams p;
    DETR_CHECK_ARG(((uintptr_t)o_lo % 16) == 0, "attention_fwd: O_lo must be 16-byte aligned");
    p.O = reinterpret_cast<__nv_bfloat16*>(o); p.o_sb = o_sb; p.o_sl = o_sl; p.O_lo = reinterpret_cast<__nv_bfloat16*>(o_lo);
    p.lse = lse; p.part = workspace;
    p.kpm = key_padding_mask; p.kpm_sb = kpm_sb; p.amask = attention_mask;
    p.B = B; p.nh = nh; p.L = L; p.S = S;
    p.scale_log2 = 1.4426950408889634f / sqrtf((float)kD);
    p.drop_thresh = dropout_threshold(dropout_p);
    p.drop_scale = (float)kDropOne / ((float)kDropOne - (float)p.drop_thresh);
    p.drop_log2_scale = log2f(p.drop_scale);
    p.seed = seed; p.seed_ptr = seed_ptr;
    p.dbg = g_fwd_dbg;
    // cudaFuncSetAttribute and the SM count are per DEVICE: remember them per device id
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    if (dev_id < 0 || dev_id >= 64) dev_id = 0;
    static bool attr_set[64] = {false};
    if (!attr_set[dev_id]) {
        cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FwdSmem::total);
        if (e != cudaSuccess) { set_error("attention_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 2; }
        attr_set[dev_id] = true;
    }
    static int sms_of[64] = {0};
    if (sms_of[dev_id] == 0 && (cudaDeviceGetAttribute(&sms_of[dev_id], cudaDevAttrMultiProcessorCount, dev_id) != cudaSuccess || sms_of[dev_id] <= 0))
        sms_of[dev_id] = 148;
    const int num_sms = sms_of[dev_id];
    const int64_t items = (int64_t)((L + kBM - 1) / kBM) * nh * B;
    const int G = (int)(items < num_sms ? items : num_sms);   // G <= items: a CTA's pair range is never shorter than one item
    cudaStream_t st = (cudaStream_t)stream;
    // decoder-sized calls (latency bound, at most 96 CTAs) may be scheduled while the projection GEMM in front of them drains
    launch_pdl_if(L <= 256, attention_fwd_kernel, dim3(G), dim3(kFwdThreads), FwdSmem::total, st, tq, tk, tv, p);
    DETR_CHECK_LAUNCH("attention_fwd");
    if (items > G) {   // only then can a boundary between two CTAs fall inside an item
        launch_pdl(attention_fwd_combine_kernel, dim3(G - 1), dim3(512), 0, st, p, G);
        DETR_CHECK_LAUNCH("attention_fwd_combine");
    }
    return 0;
}
