// Sine positional encoding and padding mask of DETR.forward in ONE launch, written directly in the token-major layout
// the encoder consumes (detr/position_encoding.py:5-97, detr/model.py:96-114).  The reference builds the coordinate grids
// with a per-image host loop (implicit .item() syncs, H2D copies) and ~15 ATen kernels; the values are:
//   ny = ceil(h / scale), nx = ceil(w / scale)                       valid window of image b on the feature map
//   gy = iy / (ny - 1) inside the window, 0 in the padding            (torch.linspace(0, 1, ny); 0 when ny == 1)
//   channel c <  F : phase = 2*pi*gy / T^(2*(c/2)/F),  sin for even c, cos for odd c
//   channel c >= F : the same with gx and c - F
//   mask[b, iy, ix] = iy >= ny && ix >= nx                             (the reference masks only the bottom-right corner)
#include "common.cuh"

namespace detr {

__global__ void __launch_bounds__(256) posenc_kernel(const int32_t* __restrict__ heights, const int32_t* __restrict__ widths, int B, int eh,
                                                     int ew, int scale, int F, float temperature, float* __restrict__ pos,
                                                     uint8_t* __restrict__ mask) {
    const int C = 2 * F;
    const int64_t n = (int64_t)B * eh * ew * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t r = i / C;
        const int ix = (int)(r % ew); r /= ew;
        const int iy = (int)(r % eh);
        const int b = (int)(r / eh);
        const int ny = (heights[b] + scale - 1) / scale, nx = (widths[b] + scale - 1) / scale;
        const bool inside = iy < ny && ix < nx;
        const bool is_y = c < F;
        const int cc = is_y ? c : c - F;
        const int idx = is_y ? iy : ix, cnt = is_y ? ny : nx;
        const float g = inside ? __fdiv_rn((float)idx, (float)max(cnt - 1, 1)) : 0.f;
        const float dim_t = powf(temperature, __fdiv_rn((float)(cc & ~1), (float)F));
        const float phase = __fdiv_rn(__fmul_rn(g, 6.283185307179586f), dim_t);
        pos[i] = (cc & 1) ? cosf(phase) : sinf(phase);
        if (c == 0 && mask != nullptr) mask[((int64_t)b * eh + iy) * ew + ix] = (iy >= ny && ix >= nx) ? 1 : 0;
    }
}

}  // namespace detr

extern "C" int detr_positional_encoding_f32(const int32_t* heights, const int32_t* widths, int B, int embed_h, int embed_w, int scale,
                                            int num_pos_feats, float temperature, float* pos, uint8_t* mask, void* stream) {
    DETR_CHECK_ARG(B >= 1 && embed_h >= 1 && embed_w >= 1 && scale >= 1 && num_pos_feats >= 2 && num_pos_feats % 2 == 0,
                   "positional_encoding: bad sizes B=%d H'=%d W'=%d scale=%d F=%d", B, embed_h, embed_w, scale, num_pos_feats);
    DETR_CHECK_ARG(heights != nullptr && widths != nullptr && pos != nullptr, "positional_encoding: null pointer");
    const int64_t n = (int64_t)B * embed_h * embed_w * 2 * num_pos_feats;
    int64_t grid = (n + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    detr::posenc_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(heights, widths, B, embed_h, embed_w, scale, num_pos_feats, temperature, pos, mask);
    DETR_CHECK_LAUNCH("positional_encoding");
    return 0;
}
