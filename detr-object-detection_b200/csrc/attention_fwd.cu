// Flash attention forward for DETR's head_dim = 32 (detr/model.py:317-352): TMA-fed tcgen05 with TMEM accumulators.
//
//   O[b,q,h,:] = dropout(softmax(Q K^T / sqrt(32) + mask)) V        LSE[b,h,q] saved for the backward kernels
//
// One CTA = one (batch, head, 128-query tile); 12 warps in 3 warpgroups:
//   warps 0-7   softmax.  Warp w owns TMEM lanes (query rows) 32*(w%4).. and the key columns 64*(w/4)..+63 of each
//               128x128 fp32 score tile: ONE TMEM read into 64 registers, local row max, a 64-thread named barrier
//               to exchange the max with the warp holding the other half of the row, exp2 / row sum / dropout /
//               bf16 P into SWIZZLE_128B shared memory (the K-major A operand of P V), and 16 of the 32 running
//               O columns in registers (rescaled on-line).
//   warp 8      TMA producer: Q once, then K/V tiles of 128 keys through a 3-stage ring (SWIZZLE_64B boxes).
//   warp 9      TMEM allocation + single-thread tcgen05.mma issue: S = Q K^T (M128 N128 K32) and
//               O_tile = P V (M128 N32 K128, V consumed MN-major straight from its natural [key][d] layout).
//   warps 10-11 idle (complete the third warpgroup so that setmaxnreg can move its registers to the softmax warps).
// The score columns are released right after they are copied to registers, so S_{j+1} = Q K_{j+1}^T runs under the
// exponentials of tile j.  Two CTAs are resident per SM (256 TMEM columns, ~92 KB shared memory each): with d = 32
// the kernel is bound by MUFU/issue slots, not by the tensor pipe, and needs the warps.
//
// Masking follows the reference: key_padding_mask / attention_mask entries get a huge FINITE negative score
// (detr/model.py:326-334 uses finfo.min, so a fully masked row is uniform, not NaN); keys beyond S (tile
// padding) get -inf and never contribute.
#include <math_constants.h>

#include "common.cuh"
#include "tc.cuh"

namespace detr {
using namespace tc;

constexpr int kBM = 128;          // queries per CTA
constexpr int kBN = 128;          // keys per tile
constexpr int kD = 32;            // head dim
constexpr int kStages = 3;
constexpr int kFwdThreads = 384;
constexpr uint32_t kTileBytes = kBN * kD * 2;  // 8 KB: one Q, K or V tile
constexpr uint32_t kTmemCols = 256;            // S: [0,128)  O: [128,160)
// Masked score: huge, finite, and a POWER OF TWO (-2^126) so that score*scale is exact and the fused
// multiply-add s*scale - m*scale cancels to exactly 0 when a whole row is masked (-> uniform probabilities).
constexpr float kMaskedScore = -8.507059173023462e37f;

struct AttnFwdParams {
    __nv_bfloat16* O; int64_t o_sb, o_sl;  // (B, L, nh*32): element strides of batch and row
    float* lse;                            // (B, nh, L)
    const uint8_t* kpm; int64_t kpm_sb;    // key padding mask (B, S) bytes, may be null
    const uint8_t* amask;                  // attention mask (L, S) bytes, may be null
    int B, nh, L, S;
    float scale_log2;                      // log2(e) / sqrt(32)
    uint32_t drop_thresh;                  // 0 = no dropout; a key is dropped if its 7 random bits < thresh
    float drop_scale;                      // 128 / (128 - thresh)
    float drop_log2_scale;                 // log2(drop_scale): folded into the exponent
    uint64_t seed; const uint64_t* seed_ptr; // effective seed = seed + *seed_ptr (device side: CUDA-graph replays get fresh masks)
};

struct FwdSmem {
    static constexpr uint32_t q = 0;
    static constexpr uint32_t k = q + kTileBytes;
    static constexpr uint32_t v = k + kStages * kTileBytes;
    static constexpr uint32_t p = 57344;                          // 128 x 128 bf16, two 64-key blocks, SWIZZLE_128B
    static constexpr uint32_t xch = p + kBM * kBN * 2;            // float[2 parity][2 halves][128 rows] row-max exchange
    static constexpr uint32_t bars = xch + 2 * 2 * kBM * 4;
    static constexpr uint32_t flags = bars + 128;
};
static_assert(FwdSmem::v + kStages * kTileBytes <= FwdSmem::p && FwdSmem::p % 1024 == 0, "smem layout");

// One 64-column half of a score tile for one query row: max, exponentials, P store.  MASKED / DROP are warp-uniform.
template <bool MASKED, bool DROP>
__device__ __forceinline__ void softmax_half_tile(uint32_t (&s)[64], const AttnFwdParams& p, const uint8_t* kf /*64 flags*/,
                                                  const uint8_t* arow /*attention-mask row at this tile's first key of the half, or null*/,
                                                  int keys_left /*S - first key of the half*/, float& m_run, float& l_run, float& alpha_out,
                                                  float* xch_mine, const float* xch_other, uint32_t bar_id, uint32_t row_key, uint32_t k32_base,
                                                  uint8_t* p_row /*this row inside the half's 64-key block*/, int row7) {
    const float sc = p.scale_log2;
    if (MASKED) {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            const uint32_t f = kf[i];
            const bool am = arow != nullptr && i < keys_left && arow[i] != 0;
            float v = __uint_as_float(s[i]);
            v = (f == 1u || am) ? kMaskedScore : v;
            v = (f == 2u) ? -CUDART_INF_F : v;
            s[i] = __float_as_uint(v);
        }
    }
    float mx = fmaxf(__uint_as_float(s[0]), __uint_as_float(s[1]));
#pragma unroll
    for (int i = 2; i < 64; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(s[i]), __uint_as_float(s[i + 1])));
    // exchange with the warp that holds the other 64 columns of this row
    *xch_mine = mx;
    named_bar_sync(bar_id, 64);
    mx = fmaxf(mx, *xch_other);
    const float m_new = fmaxf(m_run, mx);                 // finite: every tile has at least one in-range key
    alpha_out = ex2((m_run - m_new) * sc);                // first tile: exp2(-inf) = 0
    const float bias = DROP ? fmaf(-m_new, sc, p.drop_log2_scale) : -m_new * sc;   // kept entries come out pre-scaled by 1/(1-p)
    const uint32_t thr4 = p.drop_thresh * 0x01010101u;
    uint32_t rng = 0;
    float rsum = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {                         // 8 values = one 16-byte chunk of the P row
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            e[i] = ex2(fmaf(__uint_as_float(s[g * 8 + i]), sc, bias));
            rsum += e[i];
        }
        uint4 w;
        w.x = pack_bf16x2(e[0], e[1]); w.y = pack_bf16x2(e[2], e[3]); w.z = pack_bf16x2(e[4], e[5]); w.w = pack_bf16x2(e[6], e[7]);
        if (DROP) {   // the row sum above is of the un-dropped probabilities; dropped entries are cleared in the packed bf16 words
            if ((g & 3) == 0) rng = dropout_group_state(row_key, k32_base + (g >> 2));
            const uint32_t t0 = dropout_quad(rng, thr4), t1 = dropout_quad(rng, thr4);
            w.x &= dropout_mask_bf16x2<0>(t0); w.y &= dropout_mask_bf16x2<1>(t0);
            w.z &= dropout_mask_bf16x2<0>(t1); w.w &= dropout_mask_bf16x2<1>(t1);
        }
        *reinterpret_cast<uint4*>(p_row + ((g ^ row7) << 4)) = w;
    }
    l_run = l_run * alpha_out + rsum;                     // (pre-scaled by 1/(1-p) when DROP; undone in the epilogue)
    m_run = m_new;
}

__global__ void __launch_bounds__(kFwdThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const AttnFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzled tiles need 1024-byte alignment
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * kBM, h = blockIdx.y, b = blockIdx.z;
    const int T = (p.S + kBN - 1) / kBN;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::bars);
    uint64_t* q_full = bars + 0;
    uint64_t* kv_full = bars + 1;              // [kStages]
    uint64_t* kv_empty = bars + 1 + kStages;   // [kStages]
    uint64_t* s_full = bars + 1 + 2 * kStages;
    uint64_t* s_empty = s_full + 1;
    uint64_t* p_full = s_full + 2;
    uint64_t* o_full = s_full + 3;
    uint64_t* o_empty = s_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 5);
    uint8_t* kflag = smem + FwdSmem::flags;    // per key: 0 normal, 1 masked (finite), 2 beyond S (-inf)

    if (tid == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < kStages; ++s) { mbar_init(kv_full + s, 1); mbar_init(kv_empty + s, 1); }
        mbar_init(s_full, 1); mbar_init(s_empty, 256); mbar_init(p_full, 256); mbar_init(o_full, 1); mbar_init(o_empty, 256);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_slot, kTmemCols);
    for (int k = tid; k < T * kBN; k += kFwdThreads)
        kflag[k] = k >= p.S ? 2 : ((p.kpm && p.kpm[b * p.kpm_sb + k]) ? 1 : 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;

    if (warp >= 8) {
        reg_dealloc<24>();
        if (warp == 8 && lane == 0) {
            // ================= TMA producer =================
            tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
            mbar_expect_tx(q_full, kTileBytes);
            tma_load_3d(smem + FwdSmem::q, &tm_q, q_full, h * kD, q0, b);
            for (int j = 0; j < T; ++j) {
                const int s = j % kStages;
                if (j >= kStages) mbar_wait_sleep(kv_empty + s, ((j / kStages) - 1) & 1);
                mbar_expect_tx(kv_full + s, 2 * kTileBytes);
                tma_load_3d(smem + FwdSmem::k + s * kTileBytes, &tm_k, kv_full + s, h * kD, j * kBN, b);
                tma_load_3d(smem + FwdSmem::v + s * kTileBytes, &tm_v, kv_full + s, h * kD, j * kBN, b);
            }
        } else if (warp == 9 && elect_one()) {
            // ================= MMA issuer =================
            constexpr uint32_t idesc_s = make_idesc_bf16(kBM, kBN, false, false);
            constexpr uint32_t idesc_o = make_idesc_bf16(kBM, kD, false, true);
            // descriptor words (tc.cuh): high word per layout, low word = (address >> 4) + constant
            constexpr uint32_t hi64 = desc_hi(512, SWZ_64B);      // Q/K/V tiles: 64-byte rows, 8-row groups 512 B apart
            constexpr uint32_t hi128 = desc_hi(1024, SWZ_128B);   // P tile: 128-byte rows, 8-row groups 1024 B apart
            const uint32_t q_lo = smem_u32(smem + FwdSmem::q) >> 4, p_lo = smem_u32(smem + FwdSmem::p) >> 4;
            auto issue_s = [&](int j) {
                const uint32_t k_lo = smem_u32(smem + FwdSmem::k + (j % kStages) * kTileBytes) >> 4;
#pragma unroll
                for (int ks = 0; ks < kD / 16; ++ks)  // K-major, 64-byte rows, SWIZZLE_64B: 32 B per 16-channel step
                    umma_bf16_lh(tmem_s, q_lo + desc_lo(ks * 32, 16), hi64, k_lo + desc_lo(ks * 32, 16), hi64, idesc_s, ks > 0);
                umma_commit(s_full);
            };
            mbar_wait_sleep(q_full, 0);
            mbar_wait_sleep(kv_full + 0, 0);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < T; ++j) {
                if (j + 1 < T) {
                    mbar_wait_sleep(kv_full + ((j + 1) % kStages), ((j + 1) / kStages) & 1);
                    mbar_wait_sleep(s_empty, j & 1);        // the softmax warps hold S_j in registers
                    tc_fence_after();
                    issue_s(j + 1);
                }
                mbar_wait_sleep(p_full, j & 1);             // P_j is in shared memory
                if (j > 0) mbar_wait_sleep(o_empty, (j - 1) & 1);  // O_{j-1} has been read out of TMEM
                tc_fence_after();
                const uint32_t v_lo = smem_u32(smem + FwdSmem::v + (j % kStages) * kTileBytes) >> 4;
#pragma unroll
                for (int ks = 0; ks < kBN / 16; ++ks) {
                    // A = P: K-major SWIZZLE_128B, 64-key blocks of 16 KB, 32 B per 16-key step inside a block
                    // B = V: MN-major (d contiguous, 64-byte rows), SWIZZLE_64B, 16 keys = 1024 B per step
                    umma_bf16_lh(tmem_o, p_lo + desc_lo((ks >> 2) * 16384 + (ks & 3) * 32, 16), hi128,
                                 v_lo + desc_lo(ks * 1024, 512), hi64, idesc_o, ks > 0);
                }
                umma_commit(o_full);
                umma_commit(kv_empty + (j % kStages));
            }
        }
    } else {
        // ================= softmax warps =================
        reg_alloc<104>();   // 384 x 80 launch registers: 128 x 24 stay with warps 8-11, 256 x 104 <= the rest
        const int wq = warp & 3, half = warp >> 2;
        const int row = wq * 32 + lane;
        const int q = q0 + row;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        const uint32_t bh = (uint32_t)(b * p.nh + h);
        const bool drop = p.drop_thresh != 0;
        const uint32_t row_key = drop ? dropout_row_key(p.seed + (p.seed_ptr ? *p.seed_ptr : 0ull), bh, (uint32_t)q) : 0u;
        const uint8_t* arow = (p.amask && q < p.L) ? p.amask + (int64_t)q * p.S : nullptr;
        uint8_t* p_row = smem + FwdSmem::p + half * 16384 + (row >> 3) * 1024 + (row & 7) * 128;
        float* xch = reinterpret_cast<float*>(smem + FwdSmem::xch);
        float m_run = -CUDART_INF_F, l_run = 0.f;
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
        uint32_t s[64];

        for (int j = 0; j < T; ++j) {
            const int key0 = j * kBN + half * 64;
            const uint8_t* kf = kflag + key0;
            uint32_t any = p.amask ? 1u : 0u;                 // warp-uniform: does this half tile need masking at all?
            {
                const uint4* kf4 = reinterpret_cast<const uint4*>(kf);
#pragma unroll
                for (int i = 0; i < 4; ++i) { const uint4 w = kf4[i]; any |= w.x | w.y | w.z | w.w; }
            }
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            tmem_ld32(tmem_s + lane_addr + half * 64, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
            tmem_ld32(tmem_s + lane_addr + half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_empty);                             // S_{j+1} may now overwrite the score columns

            if (j > 0) {   // the P buffer is free once P V of the previous tile has completed (o_full)
                mbar_wait(o_full, (j - 1) & 1);
                tc_fence_after();
            }
            float alpha;
            float* xm = xch + (j & 1) * 256 + half * 128 + row;
            const float* xo = xch + (j & 1) * 256 + (half ^ 1) * 128 + row;
            const uint32_t k4 = (uint32_t)(key0 >> 5);   // index of the first 32-key dropout group
            const uint8_t* ar = arow ? arow + key0 : nullptr;
            // previous tile's O columns (16 per thread) are loaded first so the TMEM read overlaps the row-max exchange
            uint32_t o_prev[16];
            if (j > 0) tmem_ld16(tmem_o + lane_addr + half * 16, o_prev);
            if (any) {
                if (drop) softmax_half_tile<true, true>(s, p, kf, ar, p.S - key0, m_run, l_run, alpha, xm, xo, 1 + wq, row_key, k4, p_row, row & 7);
                else softmax_half_tile<true, false>(s, p, kf, ar, p.S - key0, m_run, l_run, alpha, xm, xo, 1 + wq, row_key, k4, p_row, row & 7);
            } else {
                if (drop) softmax_half_tile<false, true>(s, p, kf, ar, p.S - key0, m_run, l_run, alpha, xm, xo, 1 + wq, row_key, k4, p_row, row & 7);
                else softmax_half_tile<false, false>(s, p, kf, ar, p.S - key0, m_run, l_run, alpha, xm, xo, 1 + wq, row_key, k4, p_row, row & 7);
            }
            fence_proxy_async_smem();
            mbar_arrive(p_full);
            if (j > 0) {
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(o_empty);
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = (acc[i] + __uint_as_float(o_prev[i])) * alpha;
            }
        }
        // ---- epilogue: last P V, normalise, store ----
        mbar_wait(o_full, (T - 1) & 1);
        tc_fence_after();
        uint32_t o_last[16];
        tmem_ld16(tmem_o + lane_addr + half * 16, o_last);
        // total row sum = sum of the two halves (same running max, same alpha history)
        float* xm = xch + (T & 1) * 256 + half * 128 + row;
        const float* xo = xch + (T & 1) * 256 + (half ^ 1) * 128 + row;
        *xm = l_run;
        named_bar_sync(1 + wq, 64);
        float l_tot = l_run + *xo;
        if (drop) l_tot *= 1.f / p.drop_scale;              // the exponentials were pre-scaled by 1/(1-p)
        tmem_ld_wait();
        tc_fence_before();
        if (q < p.L) {
            const float inv = 1.f / l_tot;
            __nv_bfloat16* dst = p.O + b * p.o_sb + (int64_t)q * p.o_sl + h * kD + half * 16;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                uint4 w;
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = (acc[g * 8 + e] + __uint_as_float(o_last[g * 8 + e])) * inv;
                w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]); w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
                reinterpret_cast<uint4*>(dst)[g] = w;
            }
            // natural-log LSE of the scaled scores: m/sqrt(d) + ln(l).  A row whose keys are all masked is flagged
            // with +inf: the backward kernels then skip it.
            if (half == 0)
                p.lse[((int64_t)b * p.nh + h) * p.L + q] =
                    m_run == kMaskedScore ? CUDART_INF_F : (m_run * p.scale_log2 + log2f(l_tot)) * 0.6931471805599453f;
        }
    }
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, kTmemCols);
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

extern "C" int detr_attention_fwd_bf16(const void* q, int64_t q_sb, int64_t q_sl, const void* k, int64_t k_sb, int64_t k_sl,
                                       const void* v, int64_t v_sb, int64_t v_sl, void* o, int64_t o_sb, int64_t o_sl,
                                       float* lse, const uint8_t* key_padding_mask, int64_t kpm_sb,
                                       const uint8_t* attention_mask, int B, int nh, int L, int S, float dropout_p,
                                       uint64_t seed, const uint64_t* seed_ptr, void* stream) {
    DETR_CHECK_ARG(B >= 1 && nh >= 1 && L >= 1 && S >= 1, "attention_fwd: bad sizes B=%d nh=%d L=%d S=%d", B, nh, L, S);
    DETR_CHECK_ARG(B <= 65535 && nh <= 65535, "attention_fwd: B and nh must fit the grid");
    DETR_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "attention_fwd: dropout_p must be in [0,1)");
    DETR_CHECK_ARG(((uintptr_t)o % 16) == 0 && (o_sb % 8) == 0 && (o_sl % 8) == 0, "attention_fwd: O must be 16-byte aligned rows");
    const int C = nh * kD;
    CUtensorMap tq, tk, tv;
    if (int rc = make_head_tile_map(&tq, q, C, L, B, q_sl, q_sb, kBM, "attention_fwd(Q)")) return rc;
    if (int rc = make_head_tile_map(&tk, k, C, S, B, k_sl, k_sb, kBN, "attention_fwd(K)")) return rc;
    if (int rc = make_head_tile_map(&tv, v, C, S, B, v_sl, v_sb, kBN, "attention_fwd(V)")) return rc;
    AttnFwdParams p;
    p.O = reinterpret_cast<__nv_bfloat16*>(o); p.o_sb = o_sb; p.o_sl = o_sl; p.lse = lse;
    p.kpm = key_padding_mask; p.kpm_sb = kpm_sb; p.amask = attention_mask;
    p.B = B; p.nh = nh; p.L = L; p.S = S;
    p.scale_log2 = 1.4426950408889634f / sqrtf((float)kD);
    p.drop_thresh = (uint32_t)lrintf(dropout_p * 128.f);
    p.drop_scale = 128.f / (128.f - (float)p.drop_thresh);
    p.drop_log2_scale = log2f(p.drop_scale);
    p.seed = seed; p.seed_ptr = seed_ptr;
    const int T = (S + kBN - 1) / kBN;
    const size_t smem = FwdSmem::flags + (size_t)T * kBN + 1024;  // +1024: manual alignment slack
    DETR_CHECK_ARG(smem <= 110 * 1024, "attention_fwd: S=%d needs %zu B of shared memory", S, smem);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
        if (e != cudaSuccess) { set_error("attention_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 2; }
        attr_set = true;
    }
    dim3 grid((L + kBM - 1) / kBM, nh, B);
    attention_fwd_kernel<<<grid, kFwdThreads, smem, (cudaStream_t)stream>>>(tq, tk, tv, p);
    DETR_CHECK_LAUNCH("attention_fwd");
    return 0;
}
