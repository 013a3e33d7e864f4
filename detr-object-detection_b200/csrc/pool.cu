// Channels-last bf16 max pooling for the ResNet stem of the path's caller (torchvision `maxpool`, kernel 3, stride 2,
// padding 1 after conv1; detr/model.py:427-438).  ATen's max_pool_{forward,backward}_nhwc run at ~10% of HBM speed on the
// (8, 64, 400, 544) stem activation (0.45 + 0.96 ms per step); these are plain HBM-bound gather kernels:
//   forward   one thread = 8 channels (16 bytes) of one output pixel; also stores the window position (0..8) of the
//             arg-max per element (first maximum in row-major window order, as ATen) in a uint8 tensor
//   backward  one thread = 8 channels of one INPUT pixel; gathers from the <= 4 windows that contain it
//             (no atomics, deterministic, fp32 accumulation)
#include "common.cuh"
#include <cuda_bf16.h>

namespace detr {

constexpr int kPoolK = 3, kPoolS = 2, kPoolP = 1;

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                          uint8_t* __restrict__ idx, int B, int H, int W, int C, int Ho, int Wo) {
    const int c8n = C >> 3;
    const int64_t n = (int64_t)B * Ho * Wo * c8n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(t % c8n);
        int64_t r = t / c8n;
        const int wo = (int)(r % Wo); r /= Wo;
        const int ho = (int)(r % Ho);
        const int b = (int)(r / Ho);
        float best[8];
        uint32_t pos[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; pos[e] = 0xFFu; }
#pragma unroll
        for (int kh = 0; kh < kPoolK; ++kh) {
            const int h = ho * kPoolS - kPoolP + kh;
            if (h < 0 || h >= H) continue;
#pragma unroll
            for (int kw = 0; kw < kPoolK; ++kw) {
                const int w = wo * kPoolS - kPoolP + kw;
                if (w < 0 || w >= W) continue;
                const uint4 u = *reinterpret_cast<const uint4*>(x + (((int64_t)b * H + h) * W + w) * C + c8 * 8);
                const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(hh[e]);
                    // first maximum wins (strict >); the first valid element always replaces the -inf / 0xFF start
                    if (f.x > best[2 * e] || pos[2 * e] == 0xFFu) { best[2 * e] = f.x; pos[2 * e] = kh * kPoolK + kw; }
                    if (f.y > best[2 * e + 1] || pos[2 * e + 1] == 0xFFu) { best[2 * e + 1] = f.y; pos[2 * e + 1] = kh * kPoolK + kw; }
                }
            }
        }
        uint4 o;
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) oh[e] = __floats2bfloat162_rn(best[2 * e], best[2 * e + 1]);
        const int64_t off = (((int64_t)b * Ho + ho) * Wo + wo) * C + c8 * 8;
        *reinterpret_cast<uint4*>(y + off) = o;
        uint2 pk;
        pk.x = pos[0] | (pos[1] << 8) | (pos[2] << 16) | (pos[3] << 24);
        pk.y = pos[4] | (pos[5] << 8) | (pos[6] << 16) | (pos[7] << 24);
        *reinterpret_cast<uint2*>(idx + off) = pk;
    }
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx,
                                                          __nv_bfloat16* __restrict__ dx, int B, int H, int W, int C, int Ho, int Wo) {
    const int c8n = C >> 3;
    const int64_t n = (int64_t)B * H * W * c8n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(t % c8n);
        int64_t r = t / c8n;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const int b = (int)(r / H);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        // windows (ho, wo) with ho*S - P <= h <= ho*S - P + K - 1
        const int ho_lo = max(0, (h + kPoolP - kPoolK + kPoolS) / kPoolS), ho_hi = min(Ho - 1, (h + kPoolP) / kPoolS);
        const int wo_lo = max(0, (w + kPoolP - kPoolK + kPoolS) / kPoolS), wo_hi = min(Wo - 1, (w + kPoolP) / kPoolS);
        for (int ho = ho_lo; ho <= ho_hi; ++ho) {
            const int kh = h - (ho * kPoolS - kPoolP);
            for (int wo = wo_lo; wo <= wo_hi; ++wo) {
                const uint32_t mine = (uint32_t)(kh * kPoolK + (w - (wo * kPoolS - kPoolP)));
                const int64_t off = (((int64_t)b * Ho + ho) * Wo + wo) * C + c8 * 8;
                const uint2 pk = *reinterpret_cast<const uint2*>(idx + off);
                const uint4 u = *reinterpret_cast<const uint4*>(dy + off);
                const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(hh[e]);
                    const uint32_t word = e < 2 ? pk.x : pk.y;
                    const uint32_t p0 = (word >> (16 * (e & 1))) & 0xFFu, p1 = (word >> (16 * (e & 1) + 8)) & 0xFFu;
                    acc[2 * e] += p0 == mine ? f.x : 0.f;
                    acc[2 * e + 1] += p1 == mine ? f.y : 0.f;
                }
            }
        }
        uint4 o;
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) oh[e] = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
        *reinterpret_cast<uint4*>(dx + (((int64_t)b * H + h) * W + w) * C + c8 * 8) = o;
    }
}

static int pool_grid(int64_t n) {
    int64_t g = (n + 255) / 256;
    return (int)(g > 148 * 32 ? 148 * 32 : (g < 1 ? 1 : g));
}

}  // namespace detr

using namespace detr;

extern "C" int detr_maxpool3x3s2_out(int n) { return (n + 2 * kPoolP - kPoolK) / kPoolS + 1; }

extern "C" int detr_maxpool3x3s2_fwd_bf16(const void* x, void* y, uint8_t* idx, int B, int H, int W, int C, void* stream) {
    DETR_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "maxpool_fwd: need C %% 8 == 0 (B=%d H=%d W=%d C=%d)", B, H, W, C);
    DETR_CHECK_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0 && ((uintptr_t)idx % 8) == 0, "maxpool_fwd: alignment");
    const int Ho = detr_maxpool3x3s2_out(H), Wo = detr_maxpool3x3s2_out(W);
    const int64_t n = (int64_t)B * Ho * Wo * (C / 8);
    maxpool_fwd_kernel<<<pool_grid(n), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                       reinterpret_cast<__nv_bfloat16*>(y), idx, B, H, W, C, Ho, Wo);
    DETR_CHECK_LAUNCH("maxpool_fwd");
    return 0;
}

extern "C" int detr_maxpool3x3s2_bwd_bf16(const void* dy, const uint8_t* idx, void* dx, int B, int H, int W, int C, void* stream) {
    DETR_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "maxpool_bwd: need C %% 8 == 0 (B=%d H=%d W=%d C=%d)", B, H, W, C);
    DETR_CHECK_ARG(((uintptr_t)dy % 16) == 0 && ((uintptr_t)dx % 16) == 0 && ((uintptr_t)idx % 8) == 0, "maxpool_bwd: alignment");
    const int Ho = detr_maxpool3x3s2_out(H), Wo = detr_maxpool3x3s2_out(W);
    const int64_t n = (int64_t)B * H * W * (C / 8);
    maxpool_bwd_kernel<<<pool_grid(n), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(dy), idx,
                                                                       reinterpret_cast<__nv_bfloat16*>(dx), B, H, W, C, Ho, Wo);
    DETR_CHECK_LAUNCH("maxpool_bwd");
    return 0;
}

// ---- residual-gradient add fused with the preceding block's ReLU backward (caller-side glue of the ResNet harness) ----
// out = (a + b) * (x > 0): the gradient w.r.t. a bottleneck block's input is the sum of its first convolution's dgrad and the
// identity branch's gradient; when that input is the previous block's ReLU output, the previous block's threshold_backward can be
// applied in the same pass (ATen runs add and threshold_backward as two kernels: 6 tensor passes instead of 4).
namespace detr {
__global__ void __launch_bounds__(256) add_relu_mask_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                            const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t n8) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 va = reinterpret_cast<const uint4*>(a)[i], vb = reinterpret_cast<const uint4*>(b)[i], vx = reinterpret_cast<const uint4*>(x)[i];
        const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&va);
        const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&vb);
        const __nv_bfloat162* hx = reinterpret_cast<const __nv_bfloat162*>(&vx);
        uint4 o;
        __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 fa = __bfloat1622float2(ha[e]), fb = __bfloat1622float2(hb[e]), fx = __bfloat1622float2(hx[e]);
            ho[e] = __floats2bfloat162_rn(fx.x > 0.f ? fa.x + fb.x : 0.f, fx.y > 0.f ? fa.y + fb.y : 0.f);
        }
        reinterpret_cast<uint4*>(out)[i] = o;
    }
}
}  // namespace detr

/* out = (a + b) * (x > 0), bf16, n elements (n % 8 == 0), all four buffers with the same dense layout. */
extern "C" int detr_add_relu_mask_bf16(const void* a, const void* b, const void* x, void* out, long long n, void* stream) {
    DETR_CHECK_ARG(n >= 8 && n % 8 == 0, "add_relu_mask: n must be a positive multiple of 8 (n=%lld)", n);
    DETR_CHECK_ARG(((uintptr_t)a % 16) == 0 && ((uintptr_t)b % 16) == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0, "add_relu_mask: alignment");
    const int64_t n8 = n / 8;
    detr::add_relu_mask_kernel<<<detr::pool_grid(n8), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(b), reinterpret_cast<const __nv_bfloat16*>(x),
        reinterpret_cast<__nv_bfloat16*>(out), n8);
    DETR_CHECK_LAUNCH("add_relu_mask");
    return 0;
}

// ---- stem input: pad 3 + 2x2 space-to-depth + bf16 cast + channel padding, one pass (harness glue) --------------------
// out[b][i][j][c*4 + r*2 + s] = x[b][c][2i + r - 3][2j + s - 3] (0 outside the image), channels 4*Cin..Cout-1 zero.
// x: fp32, any strides (element units); out: bf16 NHWC-dense (B, Hs, Ws, Cout), Hs = (H+6)/2, Ws = (W+6)/2.
namespace detr {
__global__ void __launch_bounds__(256) stem_s2d_kernel(const float* __restrict__ x, int64_t sb, int64_t sc, int64_t sh, int64_t sw, int B, int Cin,
                                                       int H, int W, __nv_bfloat16* __restrict__ out, int Hs, int Ws, int Cout) {
    const int64_t n = (int64_t)B * Hs * Ws;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(t % Ws);
        int64_t r0 = t / Ws;
        const int i = (int)(r0 % Hs);
        const int b = (int)(r0 / Hs);
        __nv_bfloat16* o = out + t * Cout;
        for (int c8 = 0; c8 < Cout; c8 += 8) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int ch = c8 + e;
                const int c = ch >> 2, r = (ch >> 1) & 1, s = ch & 1;
                const int h = 2 * i + r - 3, w = 2 * j + s - 3;
                v[e] = (c < Cin && h >= 0 && h < H && w >= 0 && w < W) ? x[b * sb + c * sc + h * sh + w * sw] : 0.f;
            }
            uint4 u;
            __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) hh[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
            *reinterpret_cast<uint4*>(o + c8) = u;
        }
    }
}
}  // namespace detr

extern "C" int detr_stem_s2d_bf16(const float* x, long long sb, long long sc, long long sh, long long sw, int B, int Cin, int H, int W,
                                  void* out, int Cout, void* stream) {
    DETR_CHECK_ARG(B >= 1 && Cin >= 1 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "stem_s2d: need even H, W (B=%d Cin=%d H=%d W=%d)", B, Cin, H, W);
    DETR_CHECK_ARG(Cout >= 4 * Cin && Cout % 8 == 0 && ((uintptr_t)out % 16) == 0, "stem_s2d: Cout must be a multiple of 8 >= 4*Cin, out 16-byte aligned");
    const int Hs = (H + 6) / 2, Ws = (W + 6) / 2;
    const int64_t n = (int64_t)B * Hs * Ws;
    detr::stem_s2d_kernel<<<detr::pool_grid(n), 256, 0, (cudaStream_t)stream>>>(x, sb, sc, sh, sw, B, Cin, H, W, reinterpret_cast<__nv_bfloat16*>(out), Hs, Ws, Cout);
    DETR_CHECK_LAUNCH("stem_s2d");
    return 0;
}
