// Flash attention forward for DETR's head_dim = 32 (detr/model.py:317-352): TMA-fed tcgen05 with TMEM accumulators.
//
//   O[b,q,h,:] = dropout(softmax(Q K^T / sqrt(32) + mask)) V        LSE[b,h,q] saved for the backward kernels
//
// One CTA = one (batch, head, 128-query tile); 6 warps:
//   warps 0-3  softmax: thread r owns query row r (TMEM lane r): two passes over the 128x128 fp32 score tile
//              in TMEM (row max, then exp2 / row sum / dropout / bf16 P into swizzled shared memory) and the
//              running O row (32 fp32 registers), rescaled on-line.
//   warp 4     TMA producer: Q once, then K/V tiles of 128 keys through a 2-stage ring (SWIZZLE_64B boxes).
//   warp 5     TMEM allocation + single-thread tcgen05.mma issue: S = Q K^T (M128 N128 K32) and
//              O_tile = P V (M128 N32 K128, V consumed MN-major straight from its natural [key][d] layout).
// Two CTAs are resident per SM (256 TMEM columns and ~80 KB of shared memory each) so that one CTA's
// exponentials overlap the other's tensor work: with d = 32 the kernel is MUFU-bound, not tensor-bound.
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
constexpr int kStages = 2;
constexpr int kFwdThreads = 192;
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
    uint32_t drop_thresh;                  // 0 = no dropout; drop key if byte < thresh
    float drop_scale;                      // 256 / (256 - thresh)
    uint64_t seed;
};

struct FwdSmem {
    static constexpr uint32_t q = 0;
    static constexpr uint32_t k = q + kTileBytes;
    static constexpr uint32_t v = k + kStages * kTileBytes;
    static constexpr uint32_t p = v + kStages * kTileBytes;      // 128 x 128 bf16, two 64-wide K blocks, SWIZZLE_128B
    static constexpr uint32_t bars = p + kBM * kBN * 2;
    static constexpr uint32_t flags = bars + 128;
};
static_assert(FwdSmem::p % 1024 == 0, "P tile must be 1024-byte aligned for SWIZZLE_128B");

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
        mbar_init(s_full, 1); mbar_init(s_empty, 128); mbar_init(p_full, 128); mbar_init(o_full, 1); mbar_init(o_empty, 128);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, kTmemCols);
    for (int k = tid; k < T * kBN; k += kFwdThreads)
        kflag[k] = k >= p.S ? 2 : ((p.kpm && p.kpm[b * p.kpm_sb + k]) ? 1 : 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;

    if (warp == 4) {
        // ================= TMA producer =================
        if (lane == 0) {
            tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
            mbar_expect_tx(q_full, kTileBytes);
            tma_load_3d(smem + FwdSmem::q, &tm_q, q_full, h * kD, q0, b);
            for (int j = 0; j < T; ++j) {
                const int s = j % kStages;
                if (j >= kStages) mbar_wait(kv_empty + s, ((j / kStages) - 1) & 1);
                mbar_expect_tx(kv_full + s, 2 * kTileBytes);
                tma_load_3d(smem + FwdSmem::k + s * kTileBytes, &tm_k, kv_full + s, h * kD, j * kBN, b);
                tma_load_3d(smem + FwdSmem::v + s * kTileBytes, &tm_v, kv_full + s, h * kD, j * kBN, b);
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(kBM, kBN, false, false);
            constexpr uint32_t idesc_o = make_idesc_bf16(kBM, kD, false, true);
            const uint32_t sq = smem_u32(smem + FwdSmem::q), sp = smem_u32(smem + FwdSmem::p);
            auto issue_s = [&](int j) {
                const uint32_t sk = smem_u32(smem + FwdSmem::k + (j % kStages) * kTileBytes);
#pragma unroll
                for (int ks = 0; ks < kD / 16; ++ks)  // K-major, 64-byte rows, SWIZZLE_64B: 8-row groups 512 B apart
                    umma_bf16(tmem_s, make_smem_desc(sq + ks * 32, 16, 512, SWZ_64B), make_smem_desc(sk + ks * 32, 16, 512, SWZ_64B),
                              idesc_s, ks > 0);
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            mbar_wait(kv_full + 0, 0);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < T; ++j) {
                if (j + 1 < T) {
                    mbar_wait(kv_full + ((j + 1) % kStages), ((j + 1) / kStages) & 1);
                    mbar_wait(s_empty, j & 1);        // softmax has read S_j out of TMEM
                    tc_fence_after();
                    issue_s(j + 1);
                }
                mbar_wait(p_full, j & 1);             // P_j is in shared memory
                if (j > 0) mbar_wait(o_empty, (j - 1) & 1);  // O_{j-1} has been read out of TMEM
                tc_fence_after();
                const uint32_t sv = smem_u32(smem + FwdSmem::v + (j % kStages) * kTileBytes);
#pragma unroll
                for (int ks = 0; ks < kBN / 16; ++ks) {
                    // A = P: K-major SWIZZLE_128B, 64-key blocks of 16 KB, 32 B per 16-key step inside a block
                    const uint64_t da = make_smem_desc(sp + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, SWZ_128B);
                    // B = V: MN-major (d contiguous, 64-byte rows), SWIZZLE_64B, 16 keys = 1024 B per step
                    const uint64_t db = make_smem_desc(sv + ks * 1024, 512, 512, SWZ_64B);
                    umma_bf16(tmem_o, da, db, idesc_o, ks > 0);
                }
                umma_commit(o_full);
                umma_commit(kv_empty + (j % kStages));
            }
        }
    } else {
        // ================= softmax warps =================
        const int row = warp * 32 + lane;
        const int q = q0 + row;
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        const float sc = p.scale_log2;
        const uint32_t bh = (uint32_t)(b * p.nh + h);
        const uint32_t row_key = p.drop_thresh ? dropout_row_key(p.seed, bh, (uint32_t)q) : 0u;
        const uint8_t* arow = (p.amask && q < p.L) ? p.amask + (int64_t)q * p.S : nullptr;
        uint8_t* prow = smem + FwdSmem::p + (row >> 3) * 1024 + (row & 7) * 128;
        float m_run = -CUDART_INF_F, l_run = 0.f;
        float acc[kD];
#pragma unroll
        for (int i = 0; i < kD; ++i) acc[i] = 0.f;
        uint32_t r[32];

        for (int j = 0; j < T; ++j) {
            const uint8_t* kf = kflag + j * kBN;
            // does this tile need masking at all?  (common case: only the tail tile)
            uint32_t any = arow ? 1u : 0u;
            if (!any) {
                const uint4* kf4 = reinterpret_cast<const uint4*>(kf);
#pragma unroll
                for (int i = 0; i < kBN / 16; ++i) { const uint4 w = kf4[i]; any |= w.x | w.y | w.z | w.w; }
            }
            auto masked = [&](float s, int col) -> float {
                const uint8_t f = kf[col];
                if (f == 2) return -CUDART_INF_F;
                if (f == 1 || (arow && (j * kBN + col) < p.S && arow[j * kBN + col])) return kMaskedScore;
                return s;
            };
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            // ---- pass 1: row max ----
            float mx = -CUDART_INF_F;
#pragma unroll 1
            for (int c = 0; c < kBN / 32; ++c) {
                tmem_ld32(tmem_s + lane_addr + c * 32, r);
                tmem_ld_wait();
                if (any) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, masked(__uint_as_float(r[i]), c * 32 + i));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
                }
            }
            const float m_new = fmaxf(m_run, mx);
            const float alpha = ex2((m_run - m_new) * sc);  // first tile: exp2(-inf) = 0
            const float neg_m = -m_new * sc;
            // ---- fold in the previous tile's P V (also frees the P buffer and the O columns) ----
            if (j > 0) {
                mbar_wait(o_full, (j - 1) & 1);
                tc_fence_after();
                tmem_ld32(tmem_o + lane_addr, r);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(o_empty);
#pragma unroll
                for (int i = 0; i < kD; ++i) acc[i] = (acc[i] + __uint_as_float(r[i])) * alpha;
            }
            // ---- pass 2: probabilities ----
            float rsum = 0.f;
#pragma unroll 1
            for (int c = 0; c < kBN / 32; ++c) {
                tmem_ld32(tmem_s + lane_addr + c * 32, r);
                tmem_ld_wait();
                float pv[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float s = __uint_as_float(r[i]);
                    if (any) s = masked(s, c * 32 + i);
                    pv[i] = ex2(fmaf(s, sc, neg_m));
                    rsum += pv[i];
                }
                if (p.drop_thresh) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const uint32_t bits = dropout_bits4(row_key, (uint32_t)((j * kBN + c * 32) >> 2) + g);
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            pv[g * 4 + e] = ((bits >> (8 * e)) & 0xffu) < p.drop_thresh ? 0.f : pv[g * 4 + e] * p.drop_scale;
                    }
                }
                // 32 bf16 = four 16-byte chunks of this row, XOR-swizzled by (row % 8) inside the 128-byte line
                uint8_t* blk = prow + (c >> 1) * 16384;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int chunk = (c & 1) * 4 + g;
                    uint4 w;
                    w.x = pack_bf16x2(pv[g * 8 + 0], pv[g * 8 + 1]); w.y = pack_bf16x2(pv[g * 8 + 2], pv[g * 8 + 3]);
                    w.z = pack_bf16x2(pv[g * 8 + 4], pv[g * 8 + 5]); w.w = pack_bf16x2(pv[g * 8 + 6], pv[g * 8 + 7]);
                    *reinterpret_cast<uint4*>(blk + ((chunk ^ (row & 7)) << 4)) = w;
                }
            }
            tc_fence_before();
            mbar_arrive(s_empty);
            fence_proxy_async_smem();
            mbar_arrive(p_full);
            l_run = l_run * alpha + rsum;
            m_run = m_new;
        }
        // ---- epilogue: last P V, normalise, store ----
        mbar_wait(o_full, (T - 1) & 1);
        tc_fence_after();
        tmem_ld32(tmem_o + lane_addr, r);
        tmem_ld_wait();
        tc_fence_before();
        if (q < p.L) {
            const float inv = 1.f / l_run;
            __nv_bfloat16* dst = p.O + b * p.o_sb + (int64_t)q * p.o_sl + h * kD;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint4 w;
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = (acc[g * 8 + e] + __uint_as_float(r[g * 8 + e])) * inv;
                w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]); w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
                reinterpret_cast<uint4*>(dst)[g] = w;
            }
            // natural-log LSE of the scaled scores: m/sqrt(d) + ln(l)
            // a row whose keys are all masked is flagged with +inf: the backward kernels then skip it
            p.lse[((int64_t)b * p.nh + h) * p.L + q] =
                m_run == kMaskedScore ? CUDART_INF_F : (m_run * sc + log2f(l_run)) * 0.6931471805599453f;
        }
    }
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, kTmemCols);
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

}  // namespace detr

using namespace detr;

extern "C" int detr_attention_fwd_bf16(const void* q, int64_t q_sb, int64_t q_sl, const void* k, int64_t k_sb, int64_t k_sl,
                                       const void* v, int64_t v_sb, int64_t v_sl, void* o, int64_t o_sb, int64_t o_sl,
                                       float* lse, const uint8_t* key_padding_mask, int64_t kpm_sb,
                                       const uint8_t* attention_mask, int B, int nh, int L, int S, float dropout_p,
                                       uint64_t seed, void* stream) {
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
    p.drop_thresh = (uint32_t)lrintf(dropout_p * 256.f);
    p.drop_scale = 256.f / (256.f - (float)p.drop_thresh);
    p.seed = seed;
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
