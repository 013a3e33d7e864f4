// Hand-written tcgen05 GEMMs for the transformer's nn.Linear layers (detr/model.py:312-314,354 the four attention
// projections, :405-411 the FFN) with the surrounding row work fused in:
//
//   gemm_stream_kernel   C[M,N] = epilogue(A[M,K] . B[N,K]^T)       A, B via TMA (bf16), any K % 64 == 0
//                        (kBMn: B given as [K,N] row-major, i.e. C = A . B -- the dgrad form dX = dY . W with no
//                        transposed copy of W)
//   gemm_ln_kernel       C[M,N] = epilogue(LN(x)[+addend] . W[N,256]^T)   the pre-LN LayerNorm and the "+ positional /
//                        query embedding" of detr/model.py:221-224,173-182 applied to the A tile in the PROLOGUE: the
//                        128 x 256 row block is normalised once by the worker warps, written as the swizzled K-major
//                        operand (two variants: with and without the addend -- q/k use it, v does not) and stays
//                        resident in shared memory for all the N tiles of the CTA
//   gemm_wgrad_kernel    dW[N,K] = dY[M,N]^T . X[M,K] (+ db[N] = column sums of dY), both operands MN-major straight
//                        from their row-major activations, split over M with fp32 partials folded in a fixed order
//
// epilogues (one thread per output row: the accumulator row sits on that thread's TMEM lane):
//   EPI_BIAS      out = acc + bias                                                   (projections)
//   EPI_GELU      aux = bf16(acc + bias);  out = dropout(gelu_tanh(aux))             (first FFN layer, :405-408)
//   EPI_RES       out = res + dropout(acc + bias)                                    (output projection / second FFN layer + residual)
//   EPI_GELU_BWD  out = acc * dropout_mask/(1-p) * gelu'(aux)                         (backward through :406-408)
//   EPI_SIGMOID   out = sigmoid(acc + bias), fp32                                    (box head, detr/model.py:93)
//
// All kernels: 128 x 128 output tiles, one elected thread issues tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators
// in TMEM, double buffered so the epilogue of tile i overlaps the MMAs of tile i+1), operands staged by TMA into a
// SWIZZLE_128B ring with full/empty mbarriers.  Global traffic of the epilogue is TMA as well: every epilogue warp owns a
// private 8 KB swizzled staging patch (its 32 rows x 64 columns); residual / pre-activation tiles are TMA-loaded into it
// while the MMAs run, results are written over them and leave by a TMA store -- no per-thread global access, no
// cross-warp synchronisation.
#include <cuda_bf16.h>

#include <stdlib.h>

#include <mutex>
#include <set>
#include <utility>

#include "common.cuh"
#include "rowmath.cuh"
#include "tc.cuh"

namespace detr {
using namespace tc;

constexpr int kGM = 128;                       // rows of C per tile = TMEM lanes
constexpr int kGN = 128;                       // columns of C per tile = fp32 TMEM columns of one accumulator
constexpr int kGK = 64;                        // contraction elements per stage: one 128-byte swizzle row of bf16
constexpr uint32_t kOpBytes = kGM * kGK * 2;   // 16 KB: one operand block of a stage
constexpr int kEpiWarps = 8;                   // warp w: TMEM lane quarter w & 3, column half w >> 2
constexpr int kAccBufs = 2;
constexpr uint32_t kAccCols = kAccBufs * kGN;
constexpr uint32_t kBoxBytes = 32 * 128;       // one staging box: 32 rows x 128 bytes (64 bf16 or 32 fp32 columns)
constexpr uint32_t kStgBytes = 2 * kBoxBytes;  // per epilogue warp
constexpr uint32_t kStgTotal = kEpiWarps * kStgBytes;

enum { EPI_BIAS = 0, EPI_GELU = 1, EPI_RES = 2, EPI_GELU_BWD = 3, EPI_SIGMOID = 4 };

struct EpiParams {
    const float* bias;                       // fp32 [N] or null
    int M, N;
    uint32_t thr4; float scale;              // dropout: thresh * 0x00010001 (0 = off; thresh in 1/32768), 1 / keep
    uint64_t seed; const uint64_t* seed_ptr;
    int early_trigger;                       // small grid: let the next (programmatically launched) kernel be scheduled at once
};

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }
__device__ __forceinline__ uint8_t* align1024(uint8_t* p) { return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u); }

// ---- epilogue math: packed fp32x2 (two columns per FMA-pipe instruction), one MUFU tanh per element -----------------------
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float2 bf16x2_to_f2(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
// hs * (1 + tanh(u(a))) * a with hs = 0.5 [* dropout scale]: gelu_tanh(a) (detr/model.py:406), u = a * (c1 + c2 a^2)
__device__ __forceinline__ float2 gelu2(float2 a, float hs) {
    const float2 a2 = __fmul2_rn(a, a);
    const float2 u = __fmul2_rn(__ffma2_rn(a2, f2(0.0356774081f), f2(0.7978845608f)), a);
    const float2 t = make_float2(tanh_approx(u.x), tanh_approx(u.y));
    const float2 ha = __fmul2_rn(a, f2(hs));
    return __ffma2_rn(ha, t, ha);
}
// hs * d/dy [y (1 + tanh(u(y)))] = hs * (1 + t + y u'(y) (1 - t^2))
__device__ __forceinline__ float2 gelu_grad2(float2 y, float hs) {
    const float2 y2 = __fmul2_rn(y, y);
    const float2 u = __fmul2_rn(__ffma2_rn(y2, f2(0.0356774081f), f2(0.7978845608f)), y);
    const float2 t = make_float2(tanh_approx(u.x), tanh_approx(u.y));
    const float2 r = __fmul2_rn(__ffma2_rn(y2, f2(0.1070322243f), f2(0.7978845608f)), y);   // y * u'(y)
    const float2 rt = __fmul2_rn(r, t);
    const float2 q = __ffma2_rn(make_float2(-rt.x, -rt.y), t, r);                            // r (1 - t^2)
    return __ffma2_rn(__fadd2_rn(t, q), f2(hs), f2(hs));
}
// keep masks of the 8 elements of chunk idx8 (same generator and indexing as ew_keep8 / epilogue_bwd_kernel): half e of t[i] has
// bit 15 set iff element 2 i + e is kept
__device__ __forceinline__ void drop_pairs(uint32_t key, uint32_t idx8, uint32_t thr2, uint32_t (&t)[4]) {
    uint32_t st = dropout_group_state(key, idx8);
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = dropout_pair(st, thr2);
}
__device__ __forceinline__ void mask_packed8(uint32_t (&w)[4], const uint32_t (&t)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] &= dropout_mask_bf16x2(t[i]);
}

// 16-byte chunk c (0..7) of row r (0..31) of a SWIZZLE_128B staging box
__device__ __forceinline__ uint4* box_chunk(uint8_t* box, int r, int c) {
    return reinterpret_cast<uint4*>(box + r * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
}
__device__ __forceinline__ uint4 pack8(const float* v) {
    return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// 32 columns (`half` = which 32 of the warp's 64) of this thread's row r: accumulators -> staging patch.
// bf16 tensors use one box per 64 columns (chunk = 8 columns), fp32 tensors one box per 32 columns (chunk = 4 columns).
template <int EPI, typename TO>
__device__ __forceinline__ void epi_chunk32(const EpiParams& e, uint32_t key, int m, int n0, int r, int half, uint8_t* stg,
                                            const uint32_t (&v)[32]) {
    constexpr bool kOutF32 = sizeof(TO) == 4;
    const bool drop = e.thr4 != 0;
    const uint32_t idx8 = (uint32_t)m * (uint32_t)(e.N >> 3) + (uint32_t)(n0 >> 3);   // dropout chunk index of the first 8 columns
    float x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(v[i]);
    if (EPI != EPI_GELU_BWD && e.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n0) + j);
            const float2 lo = __fadd2_rn(make_float2(x[4 * j], x[4 * j + 1]), make_float2(b.x, b.y));
            const float2 hi = __fadd2_rn(make_float2(x[4 * j + 2], x[4 * j + 3]), make_float2(b.z, b.w));
            x[4 * j] = lo.x; x[4 * j + 1] = lo.y; x[4 * j + 2] = hi.x; x[4 * j + 3] = hi.y;
        }
    }
    if (EPI == EPI_GELU) {
        // the pre-activation is kept in bf16 for the backward pass (box 0); the activation is computed from the rounded value
        // so that forward and backward see the same function (box 1).  The dropout scale rides in the GELU's factor 0.5.
        const float hs = drop ? 0.5f * e.scale : 0.5f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 wy = pack8(x + 8 * j);
            *box_chunk(stg, r, half * 4 + j) = wy;
            const uint32_t wy4[4] = {wy.x, wy.y, wy.z, wy.w};
            uint32_t wo[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) { const float2 o = gelu2(bf16x2_to_f2(wy4[p]), hs); wo[p] = pack_bf16x2(o.x, o.y); }
            if (drop) { uint32_t t[4]; drop_pairs(key, idx8 + j, e.thr4, t); mask_packed8(wo, t); }
            *box_chunk(stg + kBoxBytes, r, half * 4 + j) = make_uint4(wo[0], wo[1], wo[2], wo[3]);
        }
        return;
    }
    if (EPI == EPI_GELU_BWD) {   // pre-activation tile in box 1 (bf16); result (bf16) to box 0
        const float hs = drop ? 0.5f * e.scale : 0.5f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 wy = *box_chunk(stg + kBoxBytes, r, half * 4 + j);
            const uint32_t wy4[4] = {wy.x, wy.y, wy.z, wy.w};
            uint32_t wo[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const float2 g = gelu_grad2(bf16x2_to_f2(wy4[p]), hs);
                const float2 d = __fmul2_rn(make_float2(x[8 * j + 2 * p], x[8 * j + 2 * p + 1]), g);
                wo[p] = pack_bf16x2(d.x, d.y);
            }
            if (drop) { uint32_t t[4]; drop_pairs(key, idx8 + j, e.thr4, t); mask_packed8(wo, t); }
            *box_chunk(stg, r, half * 4 + j) = make_uint4(wo[0], wo[1], wo[2], wo[3]);
        }
        return;
    }
    if (EPI == EPI_SIGMOID) {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = 1.f / (1.f + expf(-x[i]));
    }
    if (EPI == EPI_RES && drop) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t t[4];
            drop_pairs(key, idx8 + j, e.thr4, t);
            float* y = x + 8 * j;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                y[2 * i] = __uint_as_float(__float_as_uint(y[2 * i] * e.scale) & dropout_mask_f32<0>(t[i]));
                y[2 * i + 1] = __uint_as_float(__float_as_uint(y[2 * i + 1] * e.scale) & dropout_mask_f32<1>(t[i]));
            }
        }
    }
    if (kOutF32) {
        uint8_t* box = stg + half * kBoxBytes;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint4* p = box_chunk(box, r, c);
            float4 o = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
            if (EPI == EPI_RES) { const float4 rr = *reinterpret_cast<const float4*>(p); o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w; }
            *reinterpret_cast<float4*>(p) = o;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4* p = box_chunk(stg, r, half * 4 + j);
            if (EPI == EPI_RES) {
                float rr[8];
                unpack8(*p, rr);
#pragma unroll
                for (int i = 0; i < 8; ++i) x[8 * j + i] += rr[i];
            }
            *p = pack8(x + 8 * j);
        }
    }
}

// The epilogue of one 128 x 128 accumulator for this warp (lane quarter q = rows, column half h).  `n_ld` counts the TMA
// loads this warp has issued on its private barrier (phase parity).
template <int EPI, typename TO>
__device__ __forceinline__ void epi_tile(const EpiParams& e, const CUtensorMap* tm_out, const CUtensorMap* tm_aux, const CUtensorMap* tm_res,
                                         uint32_t key, uint32_t tmem_acc, uint8_t* stg, uint64_t* ldbar, int& n_ld, int warp, int lane,
                                         int m0, int n0, uint64_t* acc_full, uint32_t acc_parity, uint64_t* acc_empty, int rows_valid = kGM) {
    constexpr bool kOutF32 = sizeof(TO) == 4;
    constexpr bool kLoad = EPI == EPI_RES || EPI == EPI_GELU_BWD;
    const int q = warp & 3, h = warp >> 2;
    const int mr = m0 + q * 32, nc = n0 + h * 64;
    const bool active = nc < e.N && q * 32 < rows_valid && mr < e.M;   // warp-uniform; rows beyond rows_valid belong to other CTAs
    // the previous tile's TMA stores have finished READING the patch before anything is written into it
    if (lane == 0) tma_store_wait_read0();
    __syncwarp();
    if (kLoad && active && lane == 0) {
        if (EPI == EPI_RES) {
            mbar_expect_tx(ldbar, kOutF32 ? 2 * kBoxBytes : kBoxBytes);
            tma_load_2d(stg, tm_res, ldbar, nc, mr);
            if (kOutF32) tma_load_2d(stg + kBoxBytes, tm_res, ldbar, nc + 32, mr);
        } else {
            mbar_expect_tx(ldbar, kBoxBytes);
            tma_load_2d(stg + kBoxBytes, tm_aux, ldbar, nc, mr);
        }
    }
    // this tile's 64 bias values (two 128-byte lines) are pulled into L1 while the MMAs run: the loads in epi_chunk32 then hit
    if (EPI != EPI_GELU_BWD && active && e.bias != nullptr && lane < 2)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(e.bias + nc + lane * 32));
    mbar_wait_sleep(acc_full, acc_parity);
    tc_fence_after();
    const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 64);
    const int m = mr + lane;
    // the two 32-column halves run through ONE copy of the epilogue code (rolled loop: the instruction footprint of these
    // short-lived kernels matters more than the second TMEM load being issued a little later)
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(taddr + 32 * half, v);
        tmem_ld_wait();
        if (half == 1) {   // both halves are in registers: the accumulator buffer is free
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
        if (half == 0 && active && kLoad) { mbar_wait_sleep(ldbar, n_ld & 1); ++n_ld; }
        if (active) epi_chunk32<EPI, TO>(e, key, m, nc + 32 * half, lane, half, stg, v);
    }
    if (!active) return;
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        if (EPI == EPI_GELU) {
            tma_store_2d(tm_aux, stg, nc, mr);
            tma_store_2d(tm_out, stg + kBoxBytes, nc, mr);
        } else if (kOutF32) {
            tma_store_2d(tm_out, stg, nc, mr);
            if (nc + 32 < e.N) tma_store_2d(tm_out, stg + kBoxBytes, nc + 32, mr);
        } else {
            tma_store_2d(tm_out, stg, nc, mr);
        }
        tma_store_commit();
    }
}

// =====================================================================================================================
// streaming GEMM
// =====================================================================================================================
constexpr int kStStages = 5;
constexpr int kStThreads = (kEpiWarps + 2) * 32;
constexpr uint32_t kStRing = kStStages * 2 * kOpBytes;
constexpr uint32_t kStSmem = kStRing + kStgTotal + 512 + 1024;

struct GemmParams {
    EpiParams e;
    int m_tiles, n_tiles, k_blocks;
};

template <int EPI, typename TO, bool kBMn>
__global__ void __launch_bounds__(kStThreads, 1)
gemm_stream_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_out,
                   const __grid_constant__ CUtensorMap tm_aux, const __grid_constant__ CUtensorMap tm_res, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* stg_all = smem + kStRing;
    uint64_t* full = reinterpret_cast<uint64_t*>(stg_all + kStgTotal);
    uint64_t* empty = full + kStStages;
    uint64_t* acc_full = empty + kStStages;
    uint64_t* acc_empty = acc_full + kAccBufs;
    uint64_t* ldbar = acc_empty + kAccBufs;                    // [kEpiWarps]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ldbar + kEpiWarps);
    if (tid == 0) {
        for (int s = 0; s < kStStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < kAccBufs; ++b) { mbar_init(acc_full + b, 1); mbar_init(acc_empty + b, kEpiWarps); }
        for (int w = 0; w < kEpiWarps; ++w) mbar_init(ldbar + w, 1);
        fence_barrier_init();
    }
    if (warp == kEpiWarps + 1) tmem_alloc(tmem_slot, kAccCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int tiles = p.m_tiles * p.n_tiles;
    pdl_wait();      // everything below reads what the previous kernel of the stream wrote
    if (p.e.early_trigger) pdl_trigger();

    if (warp == kEpiWarps) {
        if (lane == 0) {
            // ================= TMA producer =================
            tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_b);
            int it = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int mt = t / p.n_tiles, nt = t - mt * p.n_tiles;
                for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
                    const int s = it % kStStages;
                    if (it >= kStStages) mbar_wait_sleep(empty + s, ((it / kStStages) - 1) & 1);
                    mbar_expect_tx(full + s, 2 * kOpBytes);
                    uint8_t* a = smem + s * 2 * kOpBytes;
                    tma_load_2d(a, &tm_a, full + s, kb * kGK, mt * kGM);
                    if (!kBMn) {
                        tma_load_2d(a + kOpBytes, &tm_b, full + s, kb * kGK, nt * kGN);
                    } else {   // B is [K][N] row-major: two boxes of 64 columns x 64 contraction rows
                        tma_load_2d(a + kOpBytes, &tm_b, full + s, nt * kGN, kb * kGK);
                        tma_load_2d(a + kOpBytes + 8192, &tm_b, full + s, nt * kGN + 64, kb * kGK);
                    }
                }
            }
        }
    } else if (warp == kEpiWarps + 1) {
        if (elect_one()) {
            // ================= MMA issuer =================
            constexpr uint32_t idesc = make_idesc_bf16(kGM, kGN, false, kBMn);
            constexpr uint32_t hi = desc_hi(1024, SWZ_128B);   // 128-byte rows, 8-row groups 1024 B apart (K-major and MN-major alike)
            int it = 0, li = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++li) {
                const int buf = li % kAccBufs;
                if (li >= kAccBufs) mbar_wait_sleep(acc_empty + buf, ((li / kAccBufs) - 1) & 1);
                tc_fence_after();
                for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
                    const int s = it % kStStages;
                    mbar_wait_sleep(full + s, (it / kStStages) & 1);
                    tc_fence_after();
                    const uint32_t a_lo = smem_u32(smem + s * 2 * kOpBytes) >> 4, b_lo = a_lo + (kOpBytes >> 4);
#pragma unroll
                    for (int k = 0; k < kGK / 16; ++k) {
                        // K-major: 32 bytes per 16-element step inside the 128-byte row.  MN-major: 16 contraction rows = 2048 B per
                        // step, the second 64-column box 8192 B further (leading byte offset)
                        const uint32_t bd = kBMn ? b_lo + desc_lo(k * 2048, 8192) : b_lo + desc_lo(k * 32, 16);
                        umma_bf16_lh(tmem + buf * kGN, a_lo + desc_lo(k * 32, 16), hi, bd, hi, idesc, (kb | k) != 0);
                    }
                    umma_commit(empty + s);
                }
                umma_commit(acc_full + buf);
            }
        }
    } else {
        // ================= epilogue warps =================
        const uint32_t key = p.e.thr4 ? ew_key(p.e.seed, p.e.seed_ptr) : 0u;
        uint8_t* stg = stg_all + warp * kStgBytes;
        int li = 0, n_ld = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++li) {
            const int mt = t / p.n_tiles, nt = t - mt * p.n_tiles;
            const int buf = li % kAccBufs;
            epi_tile<EPI, TO>(p.e, &tm_out, &tm_aux, &tm_res, key, tmem + buf * kGN, stg, ldbar + warp, n_ld, warp, lane, mt * kGM, nt * kGN,
                              acc_full + buf, (li / kAccBufs) & 1, acc_empty + buf);
        }
        if (lane == 0) tma_store_wait_read0();   // shared memory must outlive the reads; the writes complete with the grid
    }
    pdl_trigger();   // a dependent launch may start its prologue while this CTA tears down (it waits for the whole grid before reading)
    tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps + 1) tmem_dealloc(tmem, kAccCols);
}

// =====================================================================================================================
// LayerNorm-prologue GEMM (K = C = 256)
// =====================================================================================================================
constexpr int kLnC = 256;
constexpr int kLnKb = kLnC / kGK;                        // 4 k-blocks: the whole row block is resident
constexpr int kLnThreads = (kEpiWarps + 2) * 32;
constexpr uint32_t kLnABytes = kLnKb * kOpBytes;         // 64 KB per A variant
// Shared memory (224 KB): A without addend | A with addend | ring of W tiles (128 rows x 64 channels) | staging | barriers.
// The GELU epilogue stages two bf16 boxes per warp (pre-activation, activation) and keeps 2 ring stages, the bias epilogue one
// box and 4 stages; a launch without the addend variant (FFN: no "+ pos") starts its ring in the unused A region: +4 stages.
template <int EPI> struct LnCfg {
    static constexpr uint32_t stg_warp = EPI == EPI_GELU ? 2 * kBoxBytes : kBoxBytes;
    static constexpr int base_stages = EPI == EPI_GELU ? 2 : 4;
    static constexpr uint32_t stg_off = 2 * kLnABytes + base_stages * kOpBytes;
    static constexpr uint32_t bars_off = stg_off + kEpiWarps * stg_warp;
};
constexpr uint32_t kLnSmem = 2 * kLnABytes + 6 * kOpBytes + 512 + 1024;
static_assert(LnCfg<EPI_GELU>::bars_off == LnCfg<EPI_BIAS>::bars_off && LnCfg<EPI_BIAS>::bars_off + 512 + 1024 == kLnSmem, "LN smem layout");

struct LnGemmParams {
    EpiParams e;
    const void* x; int64_t x_ld;
    const float* gamma; const float* beta; float eps;
    const float* addend; int64_t add_sb, add_sr; int rows_per_batch;   // addend row of flattened row m: (m / rpb) * sb + (m % rpb) * sr
    int n_pos_end;                                  // output columns < n_pos_end take LN(x) + addend, the others LN(x)
    __nv_bfloat16* a_plain; __nv_bfloat16* a_pos;   // optional [M][256] copies of the two A variants (the weight gradients need them)
    float* mean; float* rstd;                       // optional [M]
    int m_tiles, n_tiles, groups;                   // CTA = (row block, column group): groups column groups per row block
    int rows_per_cta;                               // 128, 64 or 32 real rows per row block (the MMA tile always has 128: with few rows the
                                                    // prologue -- a latency chain per round of 32 rows -- is spread over more CTAs instead)
    long long* dbg;                                 // optional clock64 timeline of CTA 0 (DETR_GEMM_TIMELINE builds), NULL in production
};

#ifndef DETR_GEMM_TIMELINE
#define LN_STAMP(slot) do {} while (0)
#else
#define LN_STAMP(slot) do { if (p.dbg != nullptr && lane == 0 && blockIdx.x == 0) p.dbg[warp * 16 + (slot)] = clock64(); } while (0)
#endif

template <typename TX> struct RawRow;
template <> struct RawRow<float> { float4 a, b; };
template <> struct RawRow<__nv_bfloat16> { uint4 a; };
__device__ __forceinline__ void raw_load(RawRow<float>& r, const float* p) { r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4); }
__device__ __forceinline__ void raw_load(RawRow<__nv_bfloat16>& r, const __nv_bfloat16* p) { r.a = *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void raw_zero(RawRow<float>& r) { r.a = make_float4(0, 0, 0, 0); r.b = r.a; }
__device__ __forceinline__ void raw_zero(RawRow<__nv_bfloat16>& r) { r.a = make_uint4(0, 0, 0, 0); }
// 8 values as four packed pairs
__device__ __forceinline__ void raw_unpack(const RawRow<float>& r, float2 (&v)[4]) {
    v[0] = make_float2(r.a.x, r.a.y); v[1] = make_float2(r.a.z, r.a.w); v[2] = make_float2(r.b.x, r.b.y); v[3] = make_float2(r.b.z, r.b.w);
}
__device__ __forceinline__ void raw_unpack(const RawRow<__nv_bfloat16>& r, float2 (&v)[4]) {
    v[0] = bf16x2_to_f2(r.a.x); v[1] = bf16x2_to_f2(r.a.y); v[2] = bf16x2_to_f2(r.a.z); v[3] = bf16x2_to_f2(r.a.w);
}
__device__ __forceinline__ uint4 pack4(const float2 (&y)[4]) {
    return make_uint4(pack_bf16x2(y[0].x, y[0].y), pack_bf16x2(y[1].x, y[1].y), pack_bf16x2(y[2].x, y[2].y), pack_bf16x2(y[3].x, y[3].y));
}

// Normalise the 128-row block `mt` into the resident A tiles.  8 warps; warp w takes rows w, w+8, .. (sixteen rows) in four
// rounds of 4 rows -- a ROLLED loop: the fully unrolled version was ~3 000 straight-line instructions that every launch
// executed exactly once, i.e. at instruction-fetch speed (measured: 18 000 cycles per row block whatever the arithmetic).
// A round: the raw rows of the NEXT round (x and addend) are requested first, then the statistics of the round's 4 rows are
// reduced together (their butterfly steps interleave), then normalise, add the addend, pack, store to shared memory.
// Branch-free: rows beyond M re-read row M-1 (finite values that no store keeps: the global copies of the operands leave by TMA
// stores of the finished tiles, which clip at M).  Two-pass variance, packed fp32x2 arithmetic.
// lane = 8 consecutive channels = one 16-byte chunk: k-block lane >> 3, chunk lane & 7 of the 128-byte swizzle row.
template <typename TX>
__device__ __forceinline__ void ln_prologue(const LnGemmParams& p, uint8_t* a_plain_s, uint8_t* a_pos_s, int mt, int warp, int lane,
                                            bool has_plain, bool has_pos, bool side) {
    const int kRows = p.rows_per_cta / kEpiWarps;           // 16 / 8 / 4 rows per warp
    constexpr int kRound = 4;
    float2 g[4], b[4];
    raw_unpack(*reinterpret_cast<const RawRow<float>*>(p.gamma + lane * 8), g);
    raw_unpack(*reinterpret_cast<const RawRow<float>*>(p.beta + lane * 8), b);
    const uint32_t kb_off = (uint32_t)(lane >> 3) * kOpBytes;
    const int c = lane & 7;
    const int m_last = p.e.M - 1, rpb = p.rows_per_batch;
    // addend row of flattened row m: m * add_sr when the addend is one dense (M, C) block (add_sb == rpb * add_sr), else the
    // broadcast form (add_sb == 0: every batch adds the same rpb rows, e.g. the decoder's query embedding): (m mod rpb) * add_sr,
    // kept incrementally -- rows advance by kEpiWarps <= rpb (the launcher guarantees one of the two forms)
    const bool a_bcast = p.add_sb == 0 && rpb < p.e.M;
    const int m_base = mt * p.rows_per_cta;
    int a_r = has_pos && a_bcast ? (m_base + warp) % rpb : 0;
    RawRow<TX> xn[kRound];
    RawRow<float> an[kRound];
    auto request = [&](int r0) {   // raw loads of the round that starts at local row r0
#pragma unroll
        for (int u = 0; u < kRound; ++u) {
            const int m = min(m_base + warp + (r0 + u) * kEpiWarps, m_last);
            raw_load(xn[u], reinterpret_cast<const TX*>(p.x) + (int64_t)m * p.x_ld + lane * 8);
            if (has_pos) {
                raw_load(an[u], p.addend + (int64_t)(a_bcast ? a_r : m) * p.add_sr + lane * 8);
                a_r += kEpiWarps;
                a_r -= a_r >= rpb ? rpb : 0;
            }
        }
    };
    request(0);
#pragma unroll 1
    for (int r0 = 0; r0 < kRows; r0 += kRound) {
        RawRow<TX> xr[kRound];
        RawRow<float> ar[kRound];
#pragma unroll
        for (int u = 0; u < kRound; ++u) { xr[u] = xn[u]; ar[u] = an[u]; }
        if (r0 + kRound < kRows) request(r0 + kRound);
        float mu[kRound], rs[kRound];
        // ---- pass 1: means ----
#pragma unroll
        for (int u = 0; u < kRound; ++u) {
            float2 v[4];
            raw_unpack(xr[u], v);
            const float2 t = __fadd2_rn(__fadd2_rn(v[0], v[1]), __fadd2_rn(v[2], v[3]));
            mu[u] = t.x + t.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < kRound; ++u) mu[u] += __shfl_xor_sync(FULL_MASK, mu[u], o);
        }
        // ---- pass 2: variances ----
#pragma unroll
        for (int u = 0; u < kRound; ++u) {
            mu[u] *= (1.f / kLnC);
            float2 v[4];
            raw_unpack(xr[u], v);
            const float2 nmu = f2(-mu[u]);
            float2 d = __fadd2_rn(v[0], nmu);
            float2 q = __fmul2_rn(d, d);
#pragma unroll
            for (int i = 1; i < 4; ++i) { d = __fadd2_rn(v[i], nmu); q = __ffma2_rn(d, d, q); }
            rs[u] = q.x + q.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < kRound; ++u) rs[u] += __shfl_xor_sync(FULL_MASK, rs[u], o);
        }
        // ---- normalise, add, pack, store ----
#pragma unroll
        for (int u = 0; u < kRound; ++u) {
            const int r = warp + (r0 + u) * kEpiWarps, m = m_base + r;
            const float rstd = rsqrtf(rs[u] * (1.f / kLnC) + p.eps);
            float2 v[4], y[4];
            raw_unpack(xr[u], v);
            const float2 nmu = f2(-mu[u]), rs2 = f2(rstd);
#pragma unroll
            for (int i = 0; i < 4; ++i) y[i] = __ffma2_rn(__fmul2_rn(__fadd2_rn(v[i], nmu), rs2), g[i], b[i]);
            const uint32_t off = kb_off + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
            if (has_plain) *reinterpret_cast<uint4*>(a_plain_s + off) = pack4(y);
            if (has_pos) {
                float2 a[4];
                raw_unpack(ar[u], a);
#pragma unroll
                for (int i = 0; i < 4; ++i) y[i] = __fadd2_rn(y[i], a[i]);
                *reinterpret_cast<uint4*>(a_pos_s + off) = pack4(y);
            }
            if (side && lane == 0 && m <= m_last) { p.mean[m] = mu[u]; p.rstd[m] = rstd; }
        }
    }
}

// grid = row blocks x column groups: CTA (mt, gi) normalises row block mt once and computes the column tiles
// [gi * n_tiles / groups, (gi + 1) * n_tiles / groups) of it; group 0 also writes the side outputs.
template <int EPI, typename TX>
__global__ void __launch_bounds__(kLnThreads, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_aux,
               const __grid_constant__ CUtensorMap tm_aplain, const __grid_constant__ CUtensorMap tm_apos, const LnGemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kMaxStages = LnCfg<EPI>::base_stages + kLnKb;
    const bool has_pos = p.n_pos_end > 0 && p.addend != nullptr, has_plain = p.n_pos_end < p.e.N || !has_pos;
    uint8_t* a_plain_s = smem;
    uint8_t* a_pos_s = smem + kLnABytes;
    const int kLnStages = has_pos ? LnCfg<EPI>::base_stages : kMaxStages;                  // (run-time: see LnCfg)
    uint8_t* ring = has_pos ? smem + 2 * kLnABytes : smem + kLnABytes;
    uint8_t* stg_all = smem + LnCfg<EPI>::stg_off;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + LnCfg<EPI>::bars_off);
    uint64_t* empty = full + kMaxStages;
    uint64_t* acc_full = empty + kMaxStages;
    uint64_t* acc_empty = acc_full + kAccBufs;
    uint64_t* a_full = acc_empty + kAccBufs;
    uint64_t* ldbar = a_full + 1;                              // [kEpiWarps] (unused: no loading epilogue here)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ldbar + kEpiWarps);
    if (tid == 0) {
        for (int s = 0; s < kMaxStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < kAccBufs; ++b) { mbar_init(acc_full + b, 1); mbar_init(acc_empty + b, kEpiWarps); }
        mbar_init(a_full, kEpiWarps * 32);
        for (int w = 0; w < kEpiWarps; ++w) mbar_init(ldbar + w, 1);
        fence_barrier_init();
    }
    if (warp == kEpiWarps + 1) tmem_alloc(tmem_slot, kAccCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    LN_STAMP(0);
    const int mt = blockIdx.x / p.groups, gi = blockIdx.x - mt * p.groups;
    const int nt0 = p.n_tiles * gi / p.groups, nt1 = p.n_tiles * (gi + 1) / p.groups;
    pdl_wait();
    if (p.e.early_trigger) pdl_trigger();
    LN_STAMP(1);

    if (warp == kEpiWarps) {
        if (lane == 0) {
            // ================= TMA producer: W tiles (column tile, k-block) =================
            tma_prefetch_desc(&tm_w);
            int it = 0;
            for (int nt = nt0; nt < nt1; ++nt) {
                for (int kb = 0; kb < kLnKb; ++kb, ++it) {
                    const int s = it % kLnStages;
                    if (it >= kLnStages) mbar_wait_sleep(empty + s, ((it / kLnStages) - 1) & 1);
                    mbar_expect_tx(full + s, kOpBytes);
                    tma_load_2d(ring + s * kOpBytes, &tm_w, full + s, kb * kGK, nt * kGN);
                }
            }
            if (gi == 0 && (p.a_plain != nullptr || p.a_pos != nullptr)) {
                // the operand copies the weight-gradient GEMMs need: the finished A tiles leave by TMA stores (4 boxes of 64
                // channels x 128 rows per variant, rows >= M clipped) -- no per-thread global store in the prologue
                mbar_wait_sleep(a_full, 0);
#pragma unroll
                for (int kb = 0; kb < kLnKb; ++kb) {
                    if (has_plain && p.a_plain != nullptr) tma_store_2d(&tm_aplain, a_plain_s + kb * kOpBytes, kb * kGK, mt * p.rows_per_cta);
                    if (has_pos && p.a_pos != nullptr) tma_store_2d(&tm_apos, a_pos_s + kb * kOpBytes, kb * kGK, mt * p.rows_per_cta);
                }
                tma_store_commit();
                tma_store_wait_read0();
            }
        }
    } else if (warp == kEpiWarps + 1) {
        if (elect_one()) {
            // ================= MMA issuer =================
            constexpr uint32_t idesc = make_idesc_bf16(kGM, kGN, false, false);
            constexpr uint32_t hi = desc_hi(1024, SWZ_128B);
            mbar_wait_sleep(a_full, 0);
            LN_STAMP(2);
            int it = 0, li = 0;
            for (int nt = nt0; nt < nt1; ++nt, ++li) {
                const int buf = li % kAccBufs;
                if (li >= kAccBufs) mbar_wait_sleep(acc_empty + buf, ((li / kAccBufs) - 1) & 1);
                tc_fence_after();
                const bool use_pos = has_pos && nt * kGN < p.n_pos_end;
                const uint32_t a_lo0 = smem_u32(use_pos ? a_pos_s : a_plain_s) >> 4;
                for (int kb = 0; kb < kLnKb; ++kb, ++it) {
                    const int s = it % kLnStages;
                    mbar_wait_sleep(full + s, (it / kLnStages) & 1);
                    tc_fence_after();
                    const uint32_t a_lo = a_lo0 + (uint32_t)kb * (kOpBytes >> 4), b_lo = smem_u32(ring + s * kOpBytes) >> 4;
#pragma unroll
                    for (int k = 0; k < kGK / 16; ++k)
                        umma_bf16_lh(tmem + buf * kGN, a_lo + desc_lo(k * 32, 16), hi, b_lo + desc_lo(k * 32, 16), hi, idesc, (kb | k) != 0);
                    umma_commit(empty + s);
                }
                umma_commit(acc_full + buf);
                if (li < 6) LN_STAMP(3 + li);
            }
        }
    } else {
        // ================= worker warps: LayerNorm prologue, then the epilogues =================
        const uint32_t key = p.e.thr4 ? ew_key(p.e.seed, p.e.seed_ptr) : 0u;
        const bool side = gi == 0 && p.mean != nullptr;
        ln_prologue<TX>(p, a_plain_s, a_pos_s, mt, warp, lane, has_plain, has_pos, side);
        LN_STAMP(2);
        fence_proxy_async_smem();
        mbar_arrive(a_full);
        uint8_t* stg = stg_all + warp * LnCfg<EPI>::stg_warp;
        int li = 0, n_ld = 0;
        for (int nt = nt0; nt < nt1; ++nt, ++li) {
            const int buf = li % kAccBufs;
            epi_tile<EPI, __nv_bfloat16>(p.e, &tm_out, &tm_aux, &tm_out, key, tmem + buf * kGN, stg, ldbar + warp, n_ld, warp, lane,
                                         mt * p.rows_per_cta, nt * kGN, acc_full + buf, (li / kAccBufs) & 1, acc_empty + buf, p.rows_per_cta);
            if (li < 6) LN_STAMP(3 + li);
        }
        LN_STAMP(9);
        if (lane == 0) tma_store_wait_read0();
        LN_STAMP(10);
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    LN_STAMP(11);
    if (warp == kEpiWarps + 1) tmem_dealloc(tmem, kAccCols);
}

// =====================================================================================================================
// weight gradient: dW[N][K] = sum_m dY[m][N]^T X[m][K], db[N] = sum_m dY[m][N]
// =====================================================================================================================
constexpr int kWgStages = 5;
constexpr int kWgThreads = (kEpiWarps + 4) * 32;   // + producer, MMA issuer, two column-sum warps
constexpr uint32_t kWgRing = kWgStages * 2 * kOpBytes;
constexpr uint32_t kWgSmem = kWgRing + kStgTotal + 256 + 1024;

struct WgradParams {
    float* dbout;        // [splits][N] or null
    int M, N, K, n_tiles, k_tiles, splits, m_blocks;
    int n_switch;        // output rows (columns of dY) >= n_switch read X from the second tensor map
};

// tm_out: the fp32 output seen as [splits * N][K] (the partial slabs; the result itself when splits == 1), box 32 x 32
__global__ void __launch_bounds__(kWgThreads, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x0,
                  const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_out, const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* stg_all = smem + kWgRing;
    uint64_t* full = reinterpret_cast<uint64_t*>(stg_all + kStgTotal);
    uint64_t* empty = full + kWgStages;
    uint64_t* acc_full = empty + kWgStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    const int split = blockIdx.x % p.splits, tile = blockIdx.x / p.splits;
    const int nt = tile / p.k_tiles, kt = tile - nt * p.k_tiles;
    const int mb0 = (int)((long long)p.m_blocks * split / p.splits), mb1 = (int)((long long)p.m_blocks * (split + 1) / p.splits);
    const bool do_db = p.dbout != nullptr && kt == 0;
    if (tid == 0) {
        for (int s = 0; s < kWgStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, do_db ? 3 : 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == kEpiWarps + 1) tmem_alloc(tmem_slot, kGN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();

    if (warp == kEpiWarps) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_dy);
            const CUtensorMap* tx = nt * kGN >= p.n_switch ? &tm_x1 : &tm_x0;
            tma_prefetch_desc(tx);
            for (int mb = mb0, it = 0; mb < mb1; ++mb, ++it) {
                const int s = it % kWgStages;
                if (it >= kWgStages) mbar_wait_sleep(empty + s, ((it / kWgStages) - 1) & 1);
                mbar_expect_tx(full + s, 2 * kOpBytes);
                uint8_t* a = smem + s * 2 * kOpBytes;
                // each operand block: two boxes of 64 contiguous columns x 64 rows of m
                tma_load_2d(a, &tm_dy, full + s, nt * kGN, mb * kGK);
                tma_load_2d(a + 8192, &tm_dy, full + s, nt * kGN + 64, mb * kGK);
                tma_load_2d(a + kOpBytes, tx, full + s, kt * kGN, mb * kGK);
                tma_load_2d(a + kOpBytes + 8192, tx, full + s, kt * kGN + 64, mb * kGK);
            }
        }
    } else if (warp == kEpiWarps + 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(kGM, kGN, true, true);
            constexpr uint32_t hi = desc_hi(1024, SWZ_128B);
            for (int mb = mb0, it = 0; mb < mb1; ++mb, ++it) {
                const int s = it % kWgStages;
                mbar_wait_sleep(full + s, (it / kWgStages) & 1);
                tc_fence_after();
                const uint32_t a_lo = smem_u32(smem + s * 2 * kOpBytes) >> 4, b_lo = a_lo + (kOpBytes >> 4);
#pragma unroll
                for (int k = 0; k < kGK / 16; ++k)
                    umma_bf16_lh(tmem, a_lo + desc_lo(k * 2048, 8192), hi, b_lo + desc_lo(k * 2048, 8192), hi, idesc, (it | k) != 0);
                umma_commit(empty + s);
            }
            umma_commit(acc_full);
        }
    } else if (warp >= kEpiWarps + 2) {
        // ================= column sums of the dY blocks (bias gradient) =================
        if (do_db) {
            const int cs = warp - (kEpiWarps + 2);                 // which 64-column box of the dY block
            float acc0 = 0.f, acc1 = 0.f;
            for (int mb = mb0, it = 0; mb < mb1; ++mb, ++it) {
                const int s = it % kWgStages;
                mbar_wait_sleep(full + s, (it / kWgStages) & 1);
                const uint8_t* a = smem + s * 2 * kOpBytes + cs * 8192;
#pragma unroll 8
                for (int r = 0; r < kGK; ++r) {
                    // row r of the box: logical 16-byte chunk lane >> 2 lives at physical chunk (lane >> 2) ^ (r & 7)
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(a + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
                    acc0 += __uint_as_float(w << 16);
                    acc1 += __uint_as_float(w & 0xffff0000u);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + s);
            }
            const int n = nt * kGN + cs * 64 + lane * 2;
            if (n < p.N) *reinterpret_cast<float2*>(p.dbout + (int64_t)split * p.N + n) = make_float2(acc0, acc1);
        }
    } else {
        // ================= epilogue: the 128 x 128 fp32 tile goes to its partial slab through the staging patch =================
        mbar_wait_sleep(acc_full, 0);
        tc_fence_after();
        const int q = warp & 3, h = warp >> 2;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 64);
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr, v0);
        tmem_ld32(taddr + 32, v1);
        tmem_ld_wait();
        uint8_t* stg = stg_all + warp * kStgBytes;
        const int n0 = nt * kGN + q * 32, k0 = kt * kGN + h * 64;
        if (n0 < p.N && k0 < p.K) {   // warp-uniform
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                *box_chunk(stg, lane, c) = make_uint4(v0[4 * c], v0[4 * c + 1], v0[4 * c + 2], v0[4 * c + 3]);
                *box_chunk(stg + kBoxBytes, lane, c) = make_uint4(v1[4 * c], v1[4 * c + 1], v1[4 * c + 2], v1[4 * c + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                // rows beyond N of this slab must not spill into the next slab: the launcher guarantees N % 32 == 0
                tma_store_2d(&tm_out, stg, k0, split * p.N + n0);
                tma_store_2d(&tm_out, stg + kBoxBytes, k0 + 32, split * p.N + n0);
                tma_store_commit();
                tma_store_wait_read0();
            }
        }
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps + 1) tmem_dealloc(tmem, kGN);
}

// out[i] = sum_s part[s][i] for the N*K weight-gradient elements, then the same for the N bias-gradient elements
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, int splits, int64_t n4, float* __restrict__ out,
                                                            const float* __restrict__ dbpart, int N, float* __restrict__ db) {
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n4) {
        float4 s = __ldcg(reinterpret_cast<const float4*>(part) + i);
        for (int k = 1; k < splits; ++k) {
            const float4 a = __ldcg(reinterpret_cast<const float4*>(part) + (int64_t)k * n4 + i);
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
        reinterpret_cast<float4*>(out)[i] = s;
    } else if (dbpart != nullptr && i - n4 < N) {
        const int c = (int)(i - n4);
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += __ldcg(dbpart + (int64_t)k * N + c);
        db[c] = s;
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn gemm_get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// row-major matrix [rows][cols] (bf16, or fp32 when f32) with row stride ld elements; box = 128 bytes of columns x box_rows rows,
// SWIZZLE_128B; out-of-range elements read as zero and are not written
static int make_map_2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool f32, const char* who) {
    EncodeTiledFn enc = gemm_get_encode();
    if (!enc) { set_error("%s: cuTensorMapEncodeTiled entry point not available", who); return 2; }
    const int es = f32 ? 4 : 2;
    if (((uintptr_t)base % 16) || ((ld * es) % 16) || rows < 1 || cols < 1) {
        set_error("%s: matrix must be 16-byte aligned with a row stride that is a multiple of 16 bytes", who);
        return 1;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r); return 2; }
    return 0;
}

static int device_sms() {
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (sms[dev] == 0 && (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms[dev] <= 0)) sms[dev] = 148;
    return sms[dev];
}

// Programmatic dependent launch of the GEMM kernels (DETR_B200_GEMM_PDL = 0 never, 1 always, 2 default).  Measured on the captured
// training step (B200): with the attribute on EVERY launch the step is 0.14-0.18 ms slower (12.76 vs 12.58 ms) whether the kernels
// trigger early or late -- every CTA needs the whole SM (225 KB of shared memory, TMEM), so behind a machine-filling grid a
// dependent grid can only sit and wait.  Mode 2 uses it where the chain is launch-latency bound: grids that leave at least half
// of the SMs free (the decoder's) are launched programmatically and trigger their dependents right after their own wait
// (transformer forward + backward at S = 850: 3.35 -> 3.30 ms).
static int gemm_pdl_mode() {   // 0 off, 1 every GEMM launch, 2 only grids that leave at least half of the SMs free (early trigger)
    static const int mode = []() { const char* e = getenv("DETR_B200_GEMM_PDL"); return e && e[0] ? atoi(e) : 2; }();
    return mode;
}
static bool gemm_pdl() { return gemm_pdl_mode() == 1; }
static bool gemm_pdl_small(int grid) { return gemm_pdl_mode() == 2 && grid * 2 <= device_sms(); }


// cudaFuncSetAttribute is per (kernel, device): remember which pairs are done (keyed by the kernel's address -- kernels
// that share a signature share the template instantiation below)
template <typename K> static int opt_in_smem(K kern, uint32_t bytes, const char* who) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    int dev = 0;
    cudaGetDevice(&dev);
    const std::pair<const void*, int> key(reinterpret_cast<const void*>(kern), dev);
    std::lock_guard<std::mutex> lock(mu);
    if (done.count(key)) return 0;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e)); return 2; }
    done.insert(key);
    return 0;
}

static int fill_epi(EpiParams& e, int M, int N, const float* bias, float dropout_p, uint64_t seed, const uint64_t* seed_ptr, const char* who) {
    if (!(dropout_p >= 0.f && dropout_p < 1.f)) { set_error("%s: dropout_p must be in [0,1)", who); return 1; }
    if (!((int64_t)M * (N / 8) < (1ll << 32))) { set_error("%s: tensor too large for the 32-bit dropout chunk counter", who); return 1; }
    const uint32_t th = dropout_threshold(dropout_p);
    e.bias = bias; e.M = M; e.N = N; e.thr4 = th * 0x00010001u; e.scale = (float)kDropOne / ((float)kDropOne - (float)th); e.seed = seed; e.seed_ptr = seed_ptr;
    return 0;
}

template <int EPI, typename TO, bool kBMn>
static int launch_stream(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tx, const CUtensorMap& tr,
                         const GemmParams& p, cudaStream_t st) {
    auto kern = gemm_stream_kernel<EPI, TO, kBMn>;
    if (int rc = opt_in_smem(kern, kStSmem, "gemm")) return rc;
    const int tiles = p.m_tiles * p.n_tiles, sms = device_sms();
    GemmParams q = p;
    q.e.early_trigger = gemm_pdl_small(tiles) ? 1 : 0;
    launch_pdl_if(gemm_pdl() || q.e.early_trigger, kern, dim3(tiles < sms ? tiles : sms), dim3(kStThreads), kStSmem, st, ta, tb, to, tx, tr, q);
    DETR_CHECK_LAUNCH("gemm");
    return 0;
}

}  // namespace detr

using namespace detr;

static long long* g_gemm_dbg = nullptr;
/* debugging aid (not part of the drop-in surface): device buffer of 10*16 int64 that receives clock64 stamps of CTA 0 */
extern "C" void detr_gemm_set_debug(long long* buf) { g_gemm_dbg = buf; }

/* dtype codes: 0 = float32, 1 = bfloat16.  epilogue: 0 bias, 1 bias + GELU(tanh) + dropout (aux receives the bf16
 * pre-activation), 2 bias + dropout + residual (out and res share out_dtype), 3 GELU backward (aux = pre-activation,
 * no bias, bf16 out), 4 sigmoid(acc + bias) (fp32 out).  b_kn = 0: b is [N][K] (nn.Linear weight, C = A B^T); 1: b is [K][N] (C = A B). */
extern "C" int detr_gemm_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, int b_kn, int M, int N, int K, int epilogue,
                              const float* bias, void* out, int out_dtype, int64_t ldo, void* aux, int64_t ld_aux, const void* res,
                              int64_t ld_res, float dropout_p, uint64_t seed, const uint64_t* seed_ptr, void* stream) {
    // N: whole 32-column boxes, except for the plain fp32 outputs of the prediction heads (N = 92 / 4): there the TMA store clips
    // the last box, the weight rows beyond N are zero-filled by the TMA load and the bias is read up to the next multiple of 64
    const bool ragged_n = N % 32 != 0;
    DETR_CHECK_ARG(M >= 1 && N >= 4 && K >= 64 && K % 64 == 0 &&
                   (!ragged_n || (N % 4 == 0 && out_dtype == 0 && !b_kn && (epilogue == EPI_BIAS || epilogue == EPI_SIGMOID))),
                   "gemm: need M >= 1, K %% 64 == 0, N %% 32 == 0 (fp32 bias / sigmoid outputs: N %% 4 == 0) (M=%d N=%d K=%d)", M, N, K);
    DETR_CHECK_ARG(epilogue >= 0 && epilogue <= 4 && (out_dtype == 0 || out_dtype == 1), "gemm: bad epilogue / dtype code");
    DETR_CHECK_ARG(epilogue != EPI_SIGMOID || (out_dtype == 0 && !b_kn), "gemm: the sigmoid epilogue writes fp32 and takes b as [N][K]");
    DETR_CHECK_ARG(out != nullptr, "gemm: out is null");
    DETR_CHECK_ARG(!bias || ((uintptr_t)bias % 16) == 0, "gemm: bias must be 16-byte aligned");
    DETR_CHECK_ARG((epilogue != EPI_GELU && epilogue != EPI_GELU_BWD) || aux != nullptr, "gemm: aux required");
    DETR_CHECK_ARG(epilogue != EPI_RES || res != nullptr, "gemm: residual required");
    DETR_CHECK_ARG((epilogue != EPI_GELU && epilogue != EPI_GELU_BWD) || out_dtype == 1, "gemm: the GELU epilogues write bf16");
    const bool f32 = out_dtype == 0;
    CUtensorMap ta, tb, to, tx, tr;
    if (int rc = make_map_2d(&ta, a, M, K, lda, kGM, false, "gemm(A)")) return rc;
    if (b_kn) { if (int rc = make_map_2d(&tb, b, K, N, ldb, kGK, false, "gemm(B[K][N])")) return rc; }
    else      { if (int rc = make_map_2d(&tb, b, N, K, ldb, kGN, false, "gemm(B[N][K])")) return rc; }
    if (int rc = make_map_2d(&to, out, M, N, ldo, 32, f32, "gemm(out)")) return rc;
    tx = to; tr = to;
    if (epilogue == EPI_GELU || epilogue == EPI_GELU_BWD) { if (int rc = make_map_2d(&tx, aux, M, N, ld_aux, 32, false, "gemm(aux)")) return rc; }
    if (epilogue == EPI_RES) { if (int rc = make_map_2d(&tr, res, M, N, ld_res, 32, f32, "gemm(res)")) return rc; }
    GemmParams p{};
    if (int rc = fill_epi(p.e, M, N, bias, dropout_p, seed, seed_ptr, "gemm")) return rc;
    p.m_tiles = (M + kGM - 1) / kGM; p.n_tiles = (N + kGN - 1) / kGN; p.k_blocks = K / kGK;
    cudaStream_t st = (cudaStream_t)stream;
#define GEMM_GO(E, T)  return b_kn ? launch_stream<E, T, true>(ta, tb, to, tx, tr, p, st) : launch_stream<E, T, false>(ta, tb, to, tx, tr, p, st)
    switch (epilogue) {
        case EPI_BIAS: if (f32) GEMM_GO(EPI_BIAS, float); else GEMM_GO(EPI_BIAS, __nv_bfloat16);
        case EPI_GELU: GEMM_GO(EPI_GELU, __nv_bfloat16);
        case EPI_RES: if (f32) GEMM_GO(EPI_RES, float); else GEMM_GO(EPI_RES, __nv_bfloat16);
        case EPI_SIGMOID: return launch_stream<EPI_SIGMOID, float, false>(ta, tb, to, tx, tr, p, st);
        default: GEMM_GO(EPI_GELU_BWD, __nv_bfloat16);
    }
#undef GEMM_GO
}

// How a LayerNorm-prologue GEMM is cut into CTAs.  A CTA normalises `rpc` real rows (32 / 64 / 96 / 128; the MMA tile always has 128
// rows, the accumulator rows beyond rpc are never stored) and computes n_tiles / groups column tiles of them.  Cost model in
// cycles per CTA, from the clock64 timeline (tools/gemm_ln_timeline.py): prologue 1 000 + 2 000 per round of 32 rows (its warps
// work through the rounds one after the other); a column tile costs its epilogue whatever rpc is (every warp owns 32 rows of
// it): ~2 200 cycles with bias, ~4 000 with GELU + dropout.  The launch takes waves x (prologue + tiles per CTA x tile).  For
// M = 6 800 this picks 96-row blocks (71 x 2 = 142 CTAs on 148 SMs instead of 54 x 2 = 108): measured 21.5 -> 19.5 us (q|k|v),
// 33.7 -> 32.7 us (FFN1 + GELU), DC5 102 -> 97 us -- the epilogue, not the partition, is what bounds these kernels.
// DETR_B200_LN_PARTITION=0: the old rule (128 rows, or 32 below a quarter wave).
static void ln_partition(int M, int n_tiles, bool gelu, int sms, int* rpc_out, int* groups_out) {
    static const bool model = []() { const char* e = getenv("DETR_B200_LN_PARTITION"); return !(e && e[0] == '0'); }();
    if (!model) {
        const int rpc = (M + 127) / 128 >= sms / 4 ? 128 : 32;
        int g = sms / ((M + rpc - 1) / rpc);
        g = g > n_tiles ? n_tiles : (g < 1 ? 1 : g);
        *rpc_out = rpc; *groups_out = g;
        return;
    }
    double best = 1e30;
    for (int rpc = 128; rpc >= 32; rpc -= 32) {
        const int m_tiles = (M + rpc - 1) / rpc;
        const double prologue = 1000.0 + 2000.0 * rpc / 32, tile = gelu ? 4000.0 : 2200.0;
        for (int g = 1; g <= n_tiles; ++g) {
            const int waves = (m_tiles * g + sms - 1) / sms, per = (n_tiles + g - 1) / g;
            const double cost = waves * (prologue + per * tile);
            if (cost < best - 1e-6) { best = cost; *rpc_out = rpc; *groups_out = g; }   // ties: larger blocks, fewer CTAs
        }
    }
}

/* Host-only query (no launch): how detr_gemm_ln_bf16 would cut an (M, N) problem into CTAs on a device with `sms` SMs -- real rows per
 * row block (32 / 64 / 96 / 128) and column groups per row block.  Lets callers and tests inspect the partition model. */
extern "C" int detr_gemm_ln_partition(int M, int N, int gelu, int sms, int* rows_per_cta, int* groups) {
    DETR_CHECK_ARG(M >= 1 && N >= 1 && sms >= 1 && rows_per_cta != nullptr && groups != nullptr, "gemm_ln_partition: bad arguments");
    ln_partition(M, (N + kGN - 1) / kGN, gelu != 0, sms, rows_per_cta, groups);
    return 0;
}

/* out[M][N] (bf16) = epilogue((LayerNorm(x) [+ addend]) . w[N][256]^T): detr/model.py:221-224,173-182 fused with the projections
 * that consume the normalised rows.  Output columns < n_pos_end are computed from LN(x) + addend, the others from LN(x).
 * addend: fp32, row of flattened row m at (m / rows_per_batch) * add_sb + (m % rows_per_batch) * add_sr.
 * a_plain / a_pos (optional, bf16 [M][256]): the two operand variants, kept for the weight gradients; mean / rstd (optional). */
extern "C" int detr_gemm_ln_bf16(const void* x, int x_dtype, int64_t ldx, const float* gamma, const float* beta, float eps,
                                 const float* addend, int64_t add_sb, int64_t add_sr, int rows_per_batch, int n_pos_end,
                                 const void* w, int64_t ldw, int M, int N, int epilogue, const float* bias, void* out, int64_t ldo,
                                 void* aux, int64_t ld_aux, void* a_plain, void* a_pos, float* mean, float* rstd, float dropout_p,
                                 uint64_t seed, const uint64_t* seed_ptr, void* stream) {
    DETR_CHECK_ARG(M >= 1 && N >= 32 && N % 32 == 0, "gemm_ln: need M >= 1, N %% 32 == 0 (M=%d N=%d)", M, N);
    DETR_CHECK_ARG(epilogue == EPI_BIAS || epilogue == EPI_GELU, "gemm_ln: epilogue must be 0 (bias) or 1 (GELU)");
    DETR_CHECK_ARG((x_dtype == 0 || x_dtype == 1) && ((uintptr_t)x % 16) == 0 && ldx % 8 == 0, "gemm_ln: x must be fp32/bf16 with 16-byte aligned rows");
    DETR_CHECK_ARG(((uintptr_t)gamma % 16) == 0 && ((uintptr_t)beta % 16) == 0, "gemm_ln: gamma / beta must be 16-byte aligned");
    DETR_CHECK_ARG(n_pos_end >= 0 && n_pos_end <= N && n_pos_end % kGN == 0, "gemm_ln: n_pos_end must be a multiple of %d", kGN);
    DETR_CHECK_ARG(n_pos_end == 0 || (addend && ((uintptr_t)addend % 16) == 0 && add_sb % 4 == 0 && add_sr % 4 == 0), "gemm_ln: addend required (16-byte aligned rows)");
    {
        const int rpb_ = rows_per_batch > 0 ? rows_per_batch : M;
        DETR_CHECK_ARG(n_pos_end == 0 || rpb_ >= M || add_sb == (int64_t)rpb_ * add_sr || (add_sb == 0 && rpb_ >= 8),
                       "gemm_ln: the addend must be one dense (M, 256) block or a broadcast (add_sb == 0) of >= 8 rows");
    }
    DETR_CHECK_ARG(out != nullptr && (epilogue != EPI_GELU || aux != nullptr), "gemm_ln: out / aux required");
    DETR_CHECK_ARG(((uintptr_t)a_plain % 16) == 0 && ((uintptr_t)a_pos % 16) == 0, "gemm_ln: operand copies must be 16-byte aligned");
    DETR_CHECK_ARG(!bias || ((uintptr_t)bias % 16) == 0, "gemm_ln: bias must be 16-byte aligned");
    CUtensorMap tw, to, tx, tap, tao;
    if (int rc = make_map_2d(&tw, w, N, kLnC, ldw, kGN, false, "gemm_ln(W)")) return rc;
    if (int rc = make_map_2d(&to, out, M, N, ldo, 32, false, "gemm_ln(out)")) return rc;
    tx = to; tap = to; tao = to;
    // (rows per row block, column groups per row block): see ln_partition
    int rpc = 128, groups = 1;
    ln_partition(M, (N + kGN - 1) / kGN, epilogue == EPI_GELU, device_sms(), &rpc, &groups);
    if (a_plain) { if (int rc = make_map_2d(&tap, a_plain, M, kLnC, kLnC, rpc, false, "gemm_ln(a_plain)")) return rc; }
    if (a_pos) { if (int rc = make_map_2d(&tao, a_pos, M, kLnC, kLnC, rpc, false, "gemm_ln(a_pos)")) return rc; }
    if (epilogue == EPI_GELU) { if (int rc = make_map_2d(&tx, aux, M, N, ld_aux, 32, false, "gemm_ln(aux)")) return rc; }
    LnGemmParams p{};
    if (int rc = fill_epi(p.e, M, N, bias, dropout_p, seed, seed_ptr, "gemm_ln")) return rc;
    p.x = x; p.x_ld = ldx; p.gamma = gamma; p.beta = beta; p.eps = eps; p.addend = n_pos_end > 0 ? addend : nullptr; p.add_sb = add_sb; p.add_sr = add_sr;
    p.rows_per_batch = rows_per_batch > 0 ? rows_per_batch : M; p.n_pos_end = n_pos_end;
    p.a_plain = reinterpret_cast<__nv_bfloat16*>(a_plain); p.a_pos = reinterpret_cast<__nv_bfloat16*>(a_pos); p.mean = mean; p.rstd = rstd;
    p.rows_per_cta = rpc;
    p.m_tiles = (M + rpc - 1) / rpc; p.n_tiles = (N + kGN - 1) / kGN;
    p.dbg = g_gemm_dbg;
    p.groups = groups;
    const dim3 grid(p.m_tiles * groups);
    cudaStream_t st = (cudaStream_t)stream;
#define LN_GO(E, TX)                                                             \
    do {                                                                         \
        auto kern = gemm_ln_kernel<E, TX>;                                       \
        if (int rc = opt_in_smem(kern, kLnSmem, "gemm_ln")) return rc;           \
        p.e.early_trigger = gemm_pdl_small((int)grid.x) ? 1 : 0;                 \
        launch_pdl_if(gemm_pdl() || p.e.early_trigger, kern, grid, dim3(kLnThreads), kLnSmem, st, tw, to, tx, tap, tao, p); \
    } while (0)
    if (epilogue == EPI_BIAS) { if (x_dtype == 0) LN_GO(EPI_BIAS, float); else LN_GO(EPI_BIAS, __nv_bfloat16); }
    else                      { if (x_dtype == 0) LN_GO(EPI_GELU, float); else LN_GO(EPI_GELU, __nv_bfloat16); }
#undef LN_GO
    DETR_CHECK_LAUNCH("gemm_ln");
    return 0;
}

// Split of the M (contraction) dimension: enough CTAs to fill the machine, but every CTA keeps at least 8 blocks of 64 rows --
// below that the partial slabs and their reduction launch cost more than the MMAs they save (M = 800: no split at all)
static int wgrad_splits(int M, int N, int K) {
    const int tiles = ((N + kGN - 1) / kGN) * ((K + kGN - 1) / kGN), m_blocks = (M + kGK - 1) / kGK;
    int s = (device_sms() + tiles - 1) / tiles;
    const int cap = m_blocks / 8;
    if (s > cap) s = cap;
    if (s > 64) s = 64;
    return s < 1 ? 1 : s;
}

extern "C" int64_t detr_gemm_wgrad_workspace_floats(int M, int N, int K) {
    const int s = wgrad_splits(M, N, K);
    return s > 1 ? (int64_t)s * ((int64_t)N * K + N) : 0;
}

/* dw[N][K] (fp32, contiguous) = dy[M][N]^T x[M][K];  db[N] (optional) = column sums of dy.  Rows n >= n_switch of dw
 * are computed from x1 instead of x0 (the fused q|k|v projection: q/k rows from LN(x)+pos, v rows from LN(x)).
 * workspace: detr_gemm_wgrad_workspace_floats(M, N, K) floats. */
extern "C" int detr_gemm_wgrad_bf16(const void* dy, int64_t ld_dy, const void* x0, int64_t ld_x0, const void* x1, int64_t ld_x1, int n_switch,
                                    int M, int N, int K, float* dw, float* db, float* workspace, void* stream) {
    DETR_CHECK_ARG(M >= 1 && N >= 64 && N % 64 == 0 && K >= 64 && K % 64 == 0, "gemm_wgrad: need N %% 64 == 0, K %% 64 == 0 (M=%d N=%d K=%d)", M, N, K);
    DETR_CHECK_ARG(dw != nullptr && ((uintptr_t)dw % 16) == 0 && ((uintptr_t)db % 8) == 0, "gemm_wgrad: dw / db alignment");
    if (x1 == nullptr || n_switch >= N) { x1 = nullptr; n_switch = N; }
    DETR_CHECK_ARG(x1 == nullptr || (n_switch >= 0 && n_switch % kGN == 0), "gemm_wgrad: n_switch must be a multiple of %d", kGN);
    const int splits = wgrad_splits(M, N, K);
    DETR_CHECK_ARG(splits == 1 || (workspace && ((uintptr_t)workspace % 16) == 0), "gemm_wgrad: workspace required");
    CUtensorMap tdy, tx0, tx1, to;
    if (int rc = make_map_2d(&tdy, dy, M, N, ld_dy, kGK, false, "gemm_wgrad(dY)")) return rc;
    if (int rc = make_map_2d(&tx0, x0, M, K, ld_x0, kGK, false, "gemm_wgrad(X)")) return rc;
    if (x1 == nullptr) tx1 = tx0;
    else if (int rc = make_map_2d(&tx1, x1, M, K, ld_x1, kGK, false, "gemm_wgrad(X1)")) return rc;
    float* slab = splits > 1 ? workspace : dw;
    if (int rc = make_map_2d(&to, slab, (int64_t)splits * N, K, K, 32, true, "gemm_wgrad(dW)")) return rc;
    WgradParams p{};
    p.M = M; p.N = N; p.K = K; p.n_tiles = (N + kGN - 1) / kGN; p.k_tiles = (K + kGN - 1) / kGN; p.splits = splits; p.m_blocks = (M + kGK - 1) / kGK;
    p.n_switch = n_switch;
    p.dbout = db ? (splits > 1 ? workspace + (int64_t)splits * N * K : db) : nullptr;
    if (int rc = opt_in_smem(gemm_wgrad_kernel, kWgSmem, "gemm_wgrad")) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    launch_pdl_if(gemm_pdl(), gemm_wgrad_kernel, dim3(p.n_tiles * p.k_tiles * splits), dim3(kWgThreads), kWgSmem, st, tdy, tx0, tx1, to, p);
    DETR_CHECK_LAUNCH("gemm_wgrad");
    if (splits > 1) {
        const int64_t n4 = (int64_t)N * K / 4, total = n4 + (db ? N : 0);
        launch_pdl_if(gemm_pdl(), wgrad_reduce_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, (const float*)workspace, splits, n4, dw,
                   (const float*)(db ? workspace + (int64_t)splits * N * K : nullptr), N, db);
        DETR_CHECK_LAUNCH("gemm_wgrad_reduce");
    }
    return 0;
}
