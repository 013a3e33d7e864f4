"""Caller of the hot path for benchmarks and end-to-end tests: a DETR-R50/R101 wrapper with the reference's
parameter names (detr/model.py:31-94) and the reference's training-step body (detr/train.py:258-267).

The backbone (torchvision ResNet + FrozenBatchNorm, cuDNN), the 1x1 input projection and the prediction heads are
OUT OF SCOPE of the B200 rewrite: they stay plain PyTorch here exactly as in the reference.
What this module adds is the glue the reference's DETR.forward does around Encoder/Decoder, with the per-image
host loops of detr/position_encoding.py:57-67 and detr/model.py:96-114 replaced by vectorised device code.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import gemm, heads
from .model import DETRConfig, Decoder, Encoder


def positional_encoding_tokens(embed_h: int, embed_w: int, heights: torch.Tensor, widths: torch.Tensor, scale: int = 32,
                               num_pos_feats: int = 128, temperature: float = 10000.0):
    """-> (pos (B, H'*W', 2F) fp32 contiguous, mask (B, H'*W') bool): detr/position_encoding.py:5-97 and
    detr/model.py:96-114 in ONE kernel (`detr_positional_encoding_f32`), already in the layout Encoder / Decoder consume
    (no per-image host loop, no `.item()` syncs, no permuted views that every LayerNorm launch would have to copy)."""
    from . import _lib
    _lib.require_cuda(heights, "positional_encoding")
    B = heights.shape[0]
    h32 = heights.to(torch.int32).contiguous()
    w32 = widths.to(torch.int32).contiguous()
    pos = torch.empty(B, embed_h * embed_w, 2 * num_pos_feats, dtype=torch.float32, device=heights.device)
    mask = torch.empty(B, embed_h * embed_w, dtype=torch.uint8, device=heights.device)
    _lib.call("detr_positional_encoding_f32", h32.data_ptr(), w32.data_ptr(), B, embed_h, embed_w, int(scale), int(num_pos_feats),
              float(temperature), pos.data_ptr(), mask.data_ptr(), _lib.stream_ptr())
    return pos, mask.view(torch.bool)


def positional_encoding_device(embed_h: int, embed_w: int, heights: torch.Tensor, widths: torch.Tensor, scale: int = 32,
                               num_pos_feats: int = 128, temperature: float = 10000.0) -> torch.Tensor:
    """(B, 2*num_pos_feats, H', W') fp32 -- the reference's `PositionalEncoding.forward` result (a view of the token-major
    tensor `positional_encoding_tokens` produces)."""
    pos, _ = positional_encoding_tokens(embed_h, embed_w, heights, widths, scale, num_pos_feats, temperature)
    return pos.permute(0, 2, 1).reshape(heights.shape[0], 2 * num_pos_feats, embed_h, embed_w)


def padding_mask_device(embed_h: int, embed_w: int, heights: torch.Tensor, widths: torch.Tensor, scale: int = 32) -> torch.Tensor:
    """(B, H', W') bool, True only on the bottom-right corner [ceil(h/s):, ceil(w/s):] -- the reference's own rule
    (detr/model.py:96-114)."""
    _, mask = positional_encoding_tokens(embed_h, embed_w, heights, widths, scale, 2)
    return mask.view(heights.shape[0], embed_h, embed_w)


class _MLP(nn.Module):
    """Box head (detr/model.py:359-392): Linear/GELU(tanh) stack, keys `net.N.*`."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers, initializer_range=0.02):
        super().__init__()
        layers = []
        for i in range(num_layers):
            layers.append(nn.Linear(input_dim if i == 0 else hidden_dim, output_dim if i == num_layers - 1 else hidden_dim))
            if i < num_layers - 1:
                layers.append(nn.GELU(approximate="tanh"))
        self.net = nn.Sequential(*layers)
        for m in self.net:
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, mean=0, std=initializer_range)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        return self.net(x)


def _conv_bn(x, conv: nn.Conv2d, bn) -> torch.Tensor:
    """conv followed by FrozenBatchNorm2d with the frozen affine folded into the convolution:
    conv(x, W) * s + t == conv(x, W * s) + t, s = gamma / sqrt(var + eps), t = beta - mean * s.
    Same parameters / buffers / state_dict keys as torchvision's modules and the same gradient w.r.t. W; it only
    removes the fp32-promoting elementwise passes FrozenBatchNorm2d makes over every activation under autocast."""
    scale, shift, _ = _bn_constants(bn)
    return F.conv2d(x, conv.weight * scale, shift, conv.stride, conv.padding, conv.dilation, conv.groups)


class _ConvBiasAct(torch.autograd.Function):
    """cuDNN fused conv + bias (+ residual) + ReLU forward (`cudnn_convolution_relu` / `_add_relu`, one kernel instead of
    conv, bias add, residual add and ReLU) with the standard convolution backward; `relu=False` (the downsample branch)
    is a plain conv + bias whose backward takes the same route, so that both gradients of a block input come out of
    `convolution_backward` in the same memory format and add with a vectorised kernel.  Library calls only (out of scope)."""

    @staticmethod
    def forward(ctx, x, w, shift, z, stride, padding, dilation, groups, relu=True):
        if not relu:
            y = F.conv2d(x, w, shift, stride, padding, dilation, groups)
        elif z is None:
            y = torch.cudnn_convolution_relu(x, w, shift, stride, padding, dilation, groups)
        else:
            y = torch.cudnn_convolution_add_relu(x, w, z, 1.0, shift, stride, padding, dilation, groups)
        if relu:
            ctx.save_for_backward(x, w, y)
        else:
            ctx.save_for_backward(x, w)
        ctx.conf = (stride, padding, dilation, groups, z is not None, relu)
        return y

    @staticmethod
    def backward(ctx, g):
        stride, padding, dilation, groups, has_z, relu = ctx.conf
        if relu:
            x, w, y = ctx.saved_tensors
            g = torch.ops.aten.threshold_backward(g, y, 0)
        else:
            x, w = ctx.saved_tensors
            g = g.contiguous(memory_format=torch.channels_last) if x.is_contiguous(memory_format=torch.channels_last) else g
        dx, dw, _ = torch.ops.aten.convolution_backward(g, x, w, None, stride, padding, dilation, False, [0, 0], groups,
                                                        [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        return dx, dw, None, (g if has_z else None), None, None, None, None, None


def _bn_constants(bn):
    """(scale (C,1,1,1) fp32, shift fp32, shift bf16) of a FrozenBatchNorm2d: its buffers never change, so the affine is
    computed once per module and device instead of with five tiny kernels per convolution per step."""
    c = bn.__dict__.get("_detr_affine")
    if c is None or c[0].device != bn.weight.device:
        with torch.no_grad():
            scale = bn.weight * (bn.running_var + bn.eps).rsqrt()
            shift = bn.bias - bn.running_mean * scale
        c = (scale.view(-1, 1, 1, 1).contiguous(), shift.contiguous(), shift.to(torch.bfloat16))
        bn.__dict__["_detr_affine"] = c   # plain attribute: not a buffer, not in the state_dict
    return c


def _conv_bn_relu(x, conv: nn.Conv2d, bn, z=None, w16=None, relu: bool = True, shift=None, no_bias: bool = False) -> torch.Tensor:
    """relu(conv_bn(x) [+ z]) through the fused cuDNN kernel (bf16, channels_last).  `w16`: the folded bf16 weight from
    `FoldedConvWeights` (one launch for the whole network); folded here with three launches when absent.  `shift`
    overrides the BN shift (the block's conv3 carries the downsample branch's shift too); `no_bias` drops it."""
    scale, _, shift_bn = _bn_constants(bn)
    shift = None if no_bias else (shift_bn if shift is None else shift)
    w = w16 if w16 is not None else (conv.weight * scale).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x = x.to(torch.bfloat16)
    with torch.autocast("cuda", enabled=False):
        return _ConvBiasAct.apply(x, w, shift, z, conv.stride, conv.padding, conv.dilation, conv.groups, relu)


def _hw_flat(t: torch.Tensor):
    """(stride_o, stride_i, stride_hw) of a 4-D weight whose (h, w) dims can be walked with one stride, else None."""
    O, I, H, W = t.shape
    so, si, sh, sw = t.stride()
    if H == 1 and W == 1:
        return so, si, 1
    if W == 1:
        return so, si, sh
    if H == 1 or sh == W * sw:
        return so, si, sw
    return None


class _MaxPool3x3s2(torch.autograd.Function):
    """torchvision's stem `maxpool` (kernel 3, stride 2, padding 1) on channels_last bf16 through the gather kernels of
    csrc/pool.cu (ATen's nhwc kernels run at a tenth of HBM speed on the 8 x 64 x 400 x 544 activation)."""

    @staticmethod
    def forward(ctx, x):
        from . import _lib
        B, C, H, W = x.shape
        Ho, Wo = _lib.load().detr_maxpool3x3s2_out(H), _lib.load().detr_maxpool3x3s2_out(W)
        y = torch.empty((B, C, Ho, Wo), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        idx = torch.empty((B, Ho, Wo, C), dtype=torch.uint8, device=x.device)
        _lib.call("detr_maxpool3x3s2_fwd_bf16", x.data_ptr(), y.data_ptr(), idx.data_ptr(), B, H, W, C, _lib.stream_ptr())
        ctx.save_for_backward(idx)
        ctx.shape = (B, C, H, W)
        return y

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        (idx,) = ctx.saved_tensors
        B, C, H, W = ctx.shape
        g = g.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dx = torch.empty((B, C, H, W), dtype=torch.bfloat16, device=g.device, memory_format=torch.channels_last)
        _lib.call("detr_maxpool3x3s2_bwd_bf16", g.data_ptr(), idx.data_ptr(), dx.data_ptr(), B, H, W, C, _lib.stream_ptr())
        return dx


def _stem_maxpool(pool: nn.MaxPool2d, x: torch.Tensor) -> torch.Tensor:
    ok = (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[1] % 8 == 0 and x.is_contiguous(memory_format=torch.channels_last)
          and pool.kernel_size in (3, (3, 3)) and pool.stride in (2, (2, 2)) and pool.padding in (1, (1, 1))
          and pool.dilation in (1, (1, 1)) and not pool.ceil_mode)
    return _MaxPool3x3s2.apply(x) if ok else pool(x)


_PREFETCH_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


def _prefetch_stream(dev: torch.device) -> "torch.cuda.Stream":
    """The per-device second stream of the FORWARD pass: work that depends on the parameters only (weight shadows, frozen-BN fold).
    Kept here, not on the modules: they stay deep-copyable / picklable."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _PREFETCH_STREAMS.get(idx)
    if st is None:
        st = _PREFETCH_STREAMS[idx] = torch.cuda.Stream(dev)
    return st


class _FoldFn(torch.autograd.Function):
    """All frozen-BN folds of the network in one launch each way: w16_k = bf16(W_k * s_k) forward (OHWI storage, i.e.
    channels_last weights for cuDNN), dW_k = fp32(dW16_k * s_k) backward."""

    @staticmethod
    def forward(ctx, pack, *weights):
        ctx.pack = pack
        pack.fold(weights)
        return tuple(pack.views)

    @staticmethod
    def backward(ctx, *grads):
        gemm._SIDE.join_now()        # the bottlenecks' weight gradients may come from the second stream (_conv_backward)
        return (None, *ctx.pack.unfold(grads))


class FoldedConvWeights:
    """bf16 BN-folded shadows of every convolution weight of a backbone in ONE flat buffer (`detr_scale_cast_multi`)."""

    def __init__(self, pairs):
        from . import _lib
        if len(pairs) > _lib.FOLD_MAX_TENSORS:
            raise ValueError(f"at most {_lib.FOLD_MAX_TENSORS} convolutions per fold pack")
        self.pairs = list(pairs)
        self.views, self.flat, self._device = [], None, None
        self.run_async = False            # set by the owner: fold() then launches on the prefetch stream and leaves an event for wait_ready()
        self._ready, self._keep = None, None   # transient (None between forwards): the objects stay deep-copyable

    def __deepcopy__(self, memo):
        """A copy starts unbuilt: the folded views become outputs of an autograd node during forward (not deep-copyable), and the
        flat buffer is rebuilt on first use anyway."""
        import copy as _copy
        new = FoldedConvWeights(_copy.deepcopy(self.pairs, memo))
        new.run_async = self.run_async
        return new

    def _build(self, device):
        total = sum(c.weight.numel() for c, _ in self.pairs)
        self.flat = torch.empty(total, dtype=torch.bfloat16, device=device)
        self.views, o = [], 0
        for c, _ in self.pairs:
            O, I, H, W = c.weight.shape
            n = O * I * H * W
            self.views.append(self.flat[o:o + n].view(O, H, W, I).permute(0, 3, 1, 2))   # logical OIHW, channels_last strides
            o += n
        self.scales = [_bn_constants(bn)[0].reshape(-1).contiguous() for _, bn in self.pairs]
        self._device = device

    def _launch(self, srcs, dsts, in_code, out_code):
        from . import _lib
        t = _lib.FoldTable()
        t.n = len(srcs)
        for k, (s_, d_) in enumerate(zip(srcs, dsts)):
            O, I, H, W = s_.shape
            t.src[k], t.dst[k], t.scale[k] = s_.data_ptr(), d_.data_ptr(), self.scales[k].data_ptr()
            t.O[k], t.I[k], t.HW[k] = O, I, H * W
            for j, v in enumerate(_hw_flat(s_)):
                t.src_stride[k][j] = v
            for j, v in enumerate(_hw_flat(d_)):
                t.dst_stride[k][j] = v
        _lib.call("detr_scale_cast_multi", ctypes_byref(t), in_code, out_code, _lib.stream_ptr())

    def fold(self, weights):
        dev = weights[0].device
        if self._device != dev:
            self._build(dev)
        srcs = [w.detach() if _hw_flat(w) is not None else w.detach().contiguous() for w in weights]
        if not self.run_async or dev.type != "cuda":
            self._launch(srcs, self.views, 0, 1)
            return
        # the fold depends on the parameters only: launched on a second stream it runs beside whatever the caller does next (the
        # ResNet stem, whose own weight sits in a pack of its own); the first consumer calls wait_ready()
        side, main = _prefetch_stream(dev), torch.cuda.current_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._launch(srcs, self.views, 0, 1)
            self._ready = torch.cuda.Event()
            self._ready.record(side)
        self._keep = srcs

    def wait_ready(self) -> None:
        if self._ready is not None:
            torch.cuda.current_stream(self._device).wait_event(self._ready)
            self._ready, self._keep = None, None

    def unfold(self, grads):
        outs, srcs = [], []
        for (c, _), g, v in zip(self.pairs, grads, self.views):
            if g is None:
                g = torch.zeros_like(v)
            if g.dtype != torch.bfloat16:
                g = g.to(torch.bfloat16)
            if _hw_flat(g) is None:
                g = g.contiguous(memory_format=torch.channels_last)
            srcs.append(g)
            outs.append(torch.empty_like(c.weight, dtype=torch.float32))
        self._launch(srcs, outs, 1, 0)
        return [o if c.weight.dtype == torch.float32 else o.to(c.weight.dtype) for o, (c, _) in zip(outs, self.pairs)]

    def __call__(self):
        """-> {conv module: folded bf16 weight}; differentiable w.r.t. the fp32 conv weights."""
        ws = _FoldFn.apply(self, *[c.weight for c, _ in self.pairs])
        return {c: w for (c, _), w in zip(self.pairs, ws)}


def ctypes_byref(obj):
    import ctypes
    return ctypes.byref(obj)


def _stem_s2d(x: torch.Tensor, conv: nn.Conv2d, w16: torch.Tensor, shift: torch.Tensor) -> torch.Tensor:
    """relu(conv1(x) + shift) of the ResNet stem (7x7, stride 2, padding 3, 3 input channels) as a 4x4 stride-1 convolution
    over the 2x2 space-to-depth image: cuDNN has no tensor-core kernel for 3 input channels (1.0 ms forward + 0.49 ms wgrad on
    the 8 x 3 x 800 x 1088 batch), the 16-channel form runs in 0.28 + 0.33 ms.
        out[o,i,j] = sum_{c,u,v} w[o,c,u,v] xp[c,2i+u,2j+v],  u = 2a+r, v = 2b+s  =>  sum_{(c,r,s),a,b} w'[o,(c,r,s),a,b] S[(c,r,s),i+a,j+b]
    with xp = x zero-padded by 3, S = pixel_unshuffle(xp, 2), w' = the 7x7 kernel zero-padded to 8x8 and regrouped.  The weight
    regrouping is differentiable torch code on the folded bf16 weight (64 x 147 values), so autograd carries dW back."""
    from . import _lib
    O, Cin = w16.shape[0], w16.shape[1]
    pad_c = (-4 * Cin) % 8
    if x.dtype == torch.float32 and not x.requires_grad:
        # one pass: pad + space-to-depth + bf16 cast + channel padding (`detr_stem_s2d_bf16`), channels_last output
        B, _, H, W = x.shape
        xs = torch.empty((B, 4 * Cin + pad_c, (H + 6) // 2, (W + 6) // 2), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
        _lib.call("detr_stem_s2d_bf16", x.data_ptr(), *x.stride(), B, Cin, H, W, xs.data_ptr(), 4 * Cin + pad_c, _lib.stream_ptr())
    else:
        xs = F.pixel_unshuffle(F.pad(x.to(torch.bfloat16), (3, 3, 3, 3)), 2)            # (B, 4*Cin, (H+6)/2, (W+6)/2)
        xs = F.pad(xs, (0, 0, 0, 0, 0, pad_c)).contiguous(memory_format=torch.channels_last)
    w = F.pad(w16, (0, 1, 0, 1)).reshape(O, Cin, 4, 2, 4, 2).permute(0, 1, 3, 5, 2, 4).reshape(O, 4 * Cin, 4, 4)
    w = F.pad(w, (0, 0, 0, 0, 0, pad_c)).contiguous(memory_format=torch.channels_last)
    with torch.autocast("cuda", enabled=False):
        return _ConvBiasAct.apply(xs, w, shift, None, (1, 1), (0, 0), (1, 1), 1, True)


def _conv_conf(conv: nn.Conv2d):
    return (tuple(conv.stride), tuple(conv.padding), tuple(conv.dilation), conv.groups)


def _conv_backward(g, inp, w, conf, need_dx: bool):
    """(dx, dw) of a convolution.  With the second stream enabled (gemm._SideStream, inside a backward pass) the weight gradient is
    launched there: the input-gradient chain of the backbone does not wait for it, and `_FoldFn.backward` -- the only consumer,
    at the very end of the pass -- joins the stream before it reads.  Library kernels either way."""
    cb = torch.ops.aten.convolution_backward
    side = gemm._SIDE.fork(g.device, (g, inp, w))
    if side is None:
        dx, dw, _ = cb(g, inp, w, None, conf[0], conf[1], conf[2], False, [0, 0], conf[3], [need_dx, True, False])
        return dx, dw
    dx = cb(g, inp, w, None, conf[0], conf[1], conf[2], False, [0, 0], conf[3], [True, False, False])[0] if need_dx else None
    with torch.cuda.stream(side):
        dw = cb(g, inp, w, None, conf[0], conf[1], conf[2], False, [0, 0], conf[3], [False, True, False])[1]
    return dx, dw


class _BottleneckFn(torch.autograd.Function):
    """One torchvision Bottleneck (frozen BN folded) as a single autograd node: three / four cuDNN fused convolutions forward,
    a hand-ordered backward in which the residual-gradient add of the block input is fused with the PREVIOUS block's ReLU
    backward (`detr_add_relu_mask_bf16`: (dx_conv1 + g_identity) * (x > 0), 4 tensor passes instead of ATen's 6).
    premask_out: this block multiplies the gradient it returns by (x > 0) -- legal when x is the previous block's ReLU output;
    premasked_in: the consumer of y already applied (y > 0) to the incoming gradient, so conv3's threshold_backward is skipped.
    Library calls + one glue kernel (out of scope of the hot path; it serves the headline step time)."""

    @staticmethod
    def forward(ctx, x, w1, w2, w3, wd, s1, s2, s3, confs, premask_out, premasked_in):
        c1, c2, c3, cd = confs
        o1 = torch.cudnn_convolution_relu(x, w1, s1, *c1)
        o2 = torch.cudnn_convolution_relu(o1, w2, s2, *c2)
        idn = x if wd is None else F.conv2d(x, wd, None, *cd)
        y = torch.cudnn_convolution_add_relu(o2, w3, idn, 1.0, s3, *c3)
        ctx.save_for_backward(x, w1, w2, w3, wd, o1, o2, y)
        ctx.confs, ctx.premask_out, ctx.premasked_in = confs, premask_out, premasked_in
        return y

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        x, w1, w2, w3, wd, o1, o2, y = ctx.saved_tensors
        c1, c2, c3, cd = ctx.confs
        tb = torch.ops.aten.threshold_backward
        cl = torch.channels_last
        g = g.contiguous(memory_format=cl)
        g3 = g if ctx.premasked_in else tb(g, y, 0)
        d2, dw3 = _conv_backward(g3, o2, w3, c3, True)
        g2 = tb(d2, o2, 0)
        d1, dw2 = _conv_backward(g2, o1, w2, c2, True)
        g1 = tb(d1, o1, 0)
        need_dx = ctx.needs_input_grad[0]
        dx1, dw1 = _conv_backward(g1, x, w1, c1, need_dx)
        dwd, b = None, g3
        if wd is not None:
            b, dwd = _conv_backward(g3, x, wd, cd, need_dx)
        dx = None
        if need_dx:
            same = (dx1.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and dx1.stride() == x.stride()
                    and b.stride() == x.stride() and x.is_contiguous(memory_format=cl) and x.numel() % 8 == 0)
            if ctx.premask_out and same:
                dx = torch.empty_like(x)
                _lib.call("detr_add_relu_mask_bf16", dx1.data_ptr(), b.data_ptr(), x.data_ptr(), dx.data_ptr(), x.numel(), _lib.stream_ptr())
            else:
                dx = dx1 + b
                if ctx.premask_out:
                    dx = tb(dx, x, 0)
        return dx, dw1, dw2, dw3, dwd, None, None, None, None, None, None


def _bottleneck_fused(blk, x, w16, premask_out: bool, premasked_in: bool):
    s1, s2 = _bn_constants(blk.bn1)[2], _bn_constants(blk.bn2)[2]
    s3 = _bn_constants(blk.bn3)[2]
    wd, cd = None, None
    if blk.downsample is not None:
        s3 = blk.__dict__.get("_detr_shift3")
        if s3 is None or s3.device != x.device:
            s3 = (_bn_constants(blk.bn3)[1] + _bn_constants(blk.downsample[1])[1]).to(torch.bfloat16)
            blk.__dict__["_detr_shift3"] = s3
        wd, cd = w16[blk.downsample[0]], _conv_conf(blk.downsample[0])
    confs = (_conv_conf(blk.conv1), _conv_conf(blk.conv2), _conv_conf(blk.conv3), cd)
    with torch.autocast("cuda", enabled=False):
        return _BottleneckFn.apply(x.to(torch.bfloat16), w16[blk.conv1], w16[blk.conv2], w16[blk.conv3], wd, s1, s2, s3, confs,
                                   premask_out, premasked_in)


def _bottleneck_forward(blk, x, fused: bool, w16=None):
    identity = x
    if fused:
        g = (lambda c: w16[c]) if w16 is not None else (lambda c: None)
        out = _conv_bn_relu(x, blk.conv1, blk.bn1, w16=g(blk.conv1))
        out = _conv_bn_relu(out, blk.conv2, blk.bn2, w16=g(blk.conv2))
        shift3 = None
        if blk.downsample is not None:
            if w16 is not None:
                # the downsample branch's BN shift rides in conv3's fused bias: a separate broadcast bias add on a
                # channels_last tensor costs ATen's strided kernel (0.26 ms on the 8 x 256 x 200 x 272 block, measured)
                identity = _conv_bn_relu(x, blk.downsample[0], blk.downsample[1], w16=g(blk.downsample[0]), relu=False, no_bias=True)
                shift3 = blk.__dict__.get("_detr_shift3")
                if shift3 is None or shift3.device != x.device:
                    shift3 = (_bn_constants(blk.bn3)[1] + _bn_constants(blk.downsample[1])[1]).to(torch.bfloat16)
                    blk.__dict__["_detr_shift3"] = shift3
            else:
                identity = _conv_bn(x, blk.downsample[0], blk.downsample[1])
        return _conv_bn_relu(out, blk.conv3, blk.bn3, z=identity, w16=g(blk.conv3), shift=shift3)
    out = F.relu(_conv_bn(x, blk.conv1, blk.bn1), inplace=True)
    out = F.relu(_conv_bn(out, blk.conv2, blk.bn2), inplace=True)
    out = _conv_bn(out, blk.conv3, blk.bn3)
    if blk.downsample is not None:
        identity = _conv_bn(x, blk.downsample[0], blk.downsample[1])
    return F.relu(out + identity, inplace=True)


class _Backbone(nn.Module):
    """torchvision ResNet C5 with FrozenBatchNorm2d, random init (no network here) -- detr/model.py:427-438.
    OUT OF SCOPE code (library convolutions); executed with the frozen BN folded into the conv weights."""

    def __init__(self, name: str, fold_bn: bool = True):
        super().__init__()
        from torchvision.models import get_model
        from torchvision.models._utils import IntermediateLayerGetter
        from torchvision.ops import FrozenBatchNorm2d
        model = get_model(name, weights=None, norm_layer=FrozenBatchNorm2d)
        self.backbone = IntermediateLayerGetter(model, return_layers={"layer4": "final_feature_map"})
        self.num_channels = 2048
        self.scale = 32
        self.fold_bn = fold_bn
        self.fuse_relu = True   # cuDNN conv+bias(+add)+ReLU epilogues; needs CUDA bf16 autocast, else plain path
        self._fold = None       # FoldedConvWeights packs, built on first fused forward
        self.use_fold_pack = True   # False: fold each weight where it is used (3 launches per convolution each way)
        self.stem_space_to_depth = True   # 7x7/s2 stem as a 4x4/s1 convolution over the space-to-depth image (tensor-core kernels)
        self.fuse_block_backward = True   # one autograd node per bottleneck (ReLU backward fused with the residual-gradient add)

    def forward(self, x):
        if not self.fold_bn:
            return self.backbone(x)["final_feature_map"]
        m = self.backbone
        fused = (self.fuse_relu and x.is_cuda and torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16)
        w16 = None
        if fused and not self.use_fold_pack:
            x = _conv_bn_relu(x.contiguous(memory_format=torch.channels_last), m.conv1, m.bn1)
        elif fused:
            if self._fold is None:
                pairs = [(m.conv1, m.bn1)]
                for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
                    for blk in layer:
                        pairs += [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2), (blk.conv3, blk.bn3)]
                        if blk.downsample is not None:
                            pairs.append((blk.downsample[0], blk.downsample[1]))
                # the stem's weight in a pack of its own (needed at once), the rest in packs of <= 64 (R101 has 104 convolutions)
                # that fold on a second stream while the stem runs
                packs = [FoldedConvWeights(pairs[:1])] + [FoldedConvWeights(pairs[i:i + 64]) for i in range(1, len(pairs), 64)]
                if os.environ.get("DETR_B200_PREFETCH_SHADOWS", "1") != "0":
                    for pk in packs[1:]:
                        pk.run_async = True
                object.__setattr__(self, "_fold", packs)
            w16 = {}
            for pack in self._fold:
                w16.update(pack())
            c1 = m.conv1
            if (self.stem_space_to_depth and tuple(c1.kernel_size) == (7, 7) and tuple(c1.stride) == (2, 2) and tuple(c1.padding) == (3, 3)
                    and c1.groups == 1 and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0):
                x = _stem_s2d(x, c1, w16[c1], _bn_constants(m.bn1)[2])
            else:
                x = _conv_bn_relu(x.contiguous(memory_format=torch.channels_last), c1, m.bn1, w16=w16[c1])
        else:
            x = F.relu(_conv_bn(x, m.conv1, m.bn1), inplace=True)
        x = _stem_maxpool(m.maxpool, x) if fused else m.maxpool(x)
        if self._fold is not None and w16 is not None:
            for pack in self._fold:
                pack.wait_ready()           # the folded weights of layer1..4 (launched on the second stream before the stem)
        blocks = [blk for layer in (m.layer1, m.layer2, m.layer3, m.layer4) for blk in layer]
        if w16 is not None and self.fuse_block_backward and all(type(b).__name__ == "Bottleneck" for b in blocks):
            for i, blk in enumerate(blocks):
                # the first block's input is the max-pool output (no ReLU to fold); the last block's output leaves the backbone
                x = _bottleneck_fused(blk, x, w16, premask_out=i > 0, premasked_in=i + 1 < len(blocks))
            return x
        for blk in blocks:
            x = _bottleneck_forward(blk, x, fused, w16)
        return x


class DetrHarness(nn.Module):
    """DETR.forward (detr/model.py:68-94) around pluggable encoder/decoder implementations."""

    def __init__(self, config: DETRConfig, encoder_fn: Optional[Callable] = None, decoder_fn: Optional[Callable] = None,
                 posenc_fn: Optional[Callable] = None):
        super().__init__()
        self.config = config
        self.backbone = _Backbone(config.backbone)
        self.input_proj = nn.Conv2d(self.backbone.num_channels, config.hidden_size, kernel_size=1)
        self.object_query_embedding = nn.Embedding(config.num_object_queries, config.hidden_size)
        self.encoder = Encoder(config)
        self.decoder = Decoder(config)
        self.class_embedding = nn.Linear(config.hidden_size, config.num_classes + 1)
        self.bbox_embedding = _MLP(config.hidden_size, config.hidden_size, 4, config.box_embedding_mlp_num_layers,
                                   config.initializer_range)
        nn.init.xavier_uniform_(self.input_proj.weight)
        nn.init.zeros_(self.input_proj.bias)
        nn.init.normal_(self.object_query_embedding.weight, mean=0.0, std=config.initializer_range)
        nn.init.xavier_uniform_(self.class_embedding.weight)
        nn.init.zeros_(self.class_embedding.bias)
        # optional functional replacements (the CPU oracle plugs in here for the reference arm of bench.py)
        self._encoder_fn = encoder_fn
        self._decoder_fn = decoder_fn
        self._posenc_fn = posenc_fn   # (H', W', heights, widths, scale, F, T) -> (pos (B,S,C), mask (B,S)); default: the CUDA kernel

    def _prefetch_shadows(self, dev: torch.device) -> None:
        """bf16 weight shadows of encoder, decoder and heads on a second stream while the ResNet runs: they depend on the parameters
        only (their consumers' `refresh` calls wait for the event instead of copying again)."""
        if not (dev.type == "cuda" and torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16
                and os.environ.get("DETR_B200_PREFETCH_SHADOWS", "1") != "0"):
            return
        packs = [getattr(self.encoder, "_shadows", None), getattr(self.decoder, "_shadows", None), heads._SHADOWS.get(self.class_embedding)]
        packs = [p for p in packs if p is not None]
        if self._encoder_fn is not None or not packs:
            return
        side = _prefetch_stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for p in packs:
                p.prefetch(dev)

    def forward(self, images: torch.Tensor, heights: torch.Tensor, widths: torch.Tensor) -> Dict[str, torch.Tensor]:
        self._prefetch_shadows(images.device)
        x = self.input_proj(self.backbone(images))
        B, C, H, W = x.shape
        posenc = self._posenc_fn or positional_encoding_tokens
        pos, mask = posenc(H, W, heights, widths, self.backbone.scale, C // 2, self.config.temperature)
        x = x.flatten(2).permute(0, 2, 1)
        query_embed = self.object_query_embedding.weight.unsqueeze(0).expand(B, -1, -1)
        if self._encoder_fn is None:
            memory = self.encoder(x, position_embedding=pos, key_padding_mask=mask)
            decoded = self.decoder(memory, position_embedding=pos, object_query_embedding=query_embed, key_padding_mask=mask)
        else:
            memory = self._encoder_fn(self.encoder, x, pos, mask)
            decoded = self._decoder_fn(self.decoder, memory, pos, query_embed, mask)
        # detr/model.py:92-93 on the GEMM kernels: fp32 logits and sigmoid boxes in the dense layout matcher and criterion read
        # (heads.predict runs the plain modules for anything outside the kernels' contract; DETR_B200_FUSED_HEADS=0 forces that)
        if os.environ.get("DETR_B200_FUSED_HEADS", "1") != "0":
            logits, boxes = heads.predict(decoded, self.class_embedding, self.bbox_embedding)
        else:
            logits, boxes = self.class_embedding(decoded), self.bbox_embedding(decoded).sigmoid()
        return {"pred_logits": logits, "pred_boxes": boxes}


def make_optimizer(model: nn.Module, lr: float = 3e-4, lr_backbone_scale: float = 0.1, weight_decay: float = 1e-4, fused: bool = True,
                   capturable: bool = False):
    """AdamW with two parameter groups, backbone at 0.1x (detr/train.py:172-182)."""
    inner = model.module if hasattr(model, "module") else model
    bb = [p for n, p in inner.named_parameters() if n.startswith("backbone.") and p.requires_grad]
    rest = [p for n, p in inner.named_parameters() if not n.startswith("backbone.") and p.requires_grad]
    # group order as the reference (detr/train.py:172-181: backbone first), so optimizer state_dicts and
    # `backbone_lr, transformer_lr = scheduler.get_last_lr()` line up
    groups = [{"params": bb, "lr": lr * lr_backbone_scale}, {"params": rest, "lr": lr}]
    return torch.optim.AdamW(groups, lr=lr, weight_decay=weight_decay, fused=fused, capturable=capturable)


def train_step(model: nn.Module, criterion: nn.Module, optimizer, batch: Dict, autocast_dtype=torch.bfloat16,
               max_grad_norm: float = 1.0) -> torch.Tensor:
    """Body of the reference's step (detr/train.py:258-267): forward + criterion under autocast, sum of the
    `loss*` entries, backward (DDP all-reduce fires here), clip, AdamW, zero_grad.  Returns the detached loss."""
    dev_type = batch["image"].device.type
    with torch.autocast(device_type=dev_type, dtype=autocast_dtype, enabled=autocast_dtype is not None):
        outputs = model(batch["image"], batch["height"], batch["width"])
    # Accelerate hands the criterion fp32 outputs (convert_outputs_to_fp32); matcher + criterion are fp32
    outputs = {k: v.float() for k, v in outputs.items()}
    losses = criterion(outputs, batch)
    loss = sum(v for k, v in losses.items() if k.startswith("loss"))
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)
    optimizer.step()
    optimizer.zero_grad(set_to_none=True)
    return loss.detach()


def synthetic_batch(batch: int, height: int = 800, width: int = 1066, num_classes: int = 91, max_gt: int = 20, seed: int = 0,
                    device: str | torch.device = "cpu", pin: bool = False) -> Dict:
    """SURVEY.md 8(d) config 1/2 input: randn image zero-padded to a multiple of 32 (detr/data.py:196-203), int32
    sizes, per-image GT lists (centres U(.2,.8), sizes U(.02,.32), XYXY) with M ~ U{1..max_gt}."""
    g = torch.Generator().manual_seed(seed)
    wp = (width + 31) // 32 * 32
    hp = (height + 31) // 32 * 32
    img = torch.zeros(batch, 3, hp, wp)
    img[:, :, :height, :width] = torch.randn(batch, 3, height, width, generator=g)
    counts = torch.randint(1, max_gt + 1, (batch,), generator=g).tolist()
    labels, boxes = [], []
    for m in counts:
        c = torch.rand(m, 2, generator=g) * 0.6 + 0.2
        s = torch.rand(m, 2, generator=g) * 0.30 + 0.02
        boxes.append(torch.cat([c - s / 2, c + s / 2], dim=1))
        labels.append(torch.randint(0, num_classes, (m,), generator=g, dtype=torch.int64))
    out = {"image": img, "height": torch.full((batch,), height, dtype=torch.int32),
           "width": torch.full((batch,), width, dtype=torch.int32), "class_idx": labels, "boxes_normalized": boxes}
    if pin:
        out = {k: (v.pin_memory() if torch.is_tensor(v) else [t.pin_memory() for t in v]) for k, v in out.items()}
    return batch_to(out, device) if str(device) != "cpu" else out


def batch_to(batch: Dict, device, non_blocking: bool = True) -> Dict:
    return {k: (v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else [t.to(device, non_blocking=non_blocking) for t in v])
            for k, v in batch.items()}


def batch_bytes(batch: Dict) -> int:
    n = 0
    for v in batch.values():
        for t in (v if isinstance(v, list) else [v]):
            n += t.numel() * t.element_size()
    return n


class FlatAdamW:
    """clip_grad_norm_ + AdamW (detr/train.py:265-267, torch.optim.AdamW semantics) on flat fp32 buffers: the parameters of
    `torch_optimizer`'s groups are re-homed into ONE buffer (each `p.data` becomes a view with its original shape and
    strides), the moments live in two more, and a step is `detr_sumsq_f32` + one `detr_adamw_clip_f32` per group instead of
    ~28 ATen launches over ~330 tensors.  Hyper-parameters are read from the torch optimizer's param_groups at every step
    (LR schedulers keep working); its own state is not used."""

    def __init__(self, torch_optimizer, device):
        self.opt = torch_optimizer
        self.params, self.ranges, o = [], [], 0
        for g in torch_optimizer.param_groups:
            ps = [q for q in g["params"] if q.requires_grad]
            n = sum(q.numel() for q in ps)
            n_pad = (n + 3) // 4 * 4                     # every group starts 16-byte aligned
            self.ranges.append((o, n))
            self.params += ps
            o += n_pad
        self.total = o
        self.flat_p = torch.zeros(o, dtype=torch.float32, device=device)
        self.flat_m, self.flat_v = torch.zeros_like(self.flat_p), torch.zeros_like(self.flat_p)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.grad_views, self.offsets, o = [], [], 0
        with torch.no_grad():
            for g, (start, n) in zip(torch_optimizer.param_groups, self.ranges):
                o = start
                for q in (q for q in g["params"] if q.requires_grad):
                    view = torch.as_strided(self.flat_p, q.shape, q.stride(), o)
                    view.copy_(q)
                    q.data = view                                  # the parameter now lives in the flat buffer
                    self.grad_views.append(torch.as_strided(self.flat_g, q.shape, q.stride(), o))
                    o += q.numel()
        from . import _lib
        self.step_t = torch.zeros(1, dtype=torch.float32, device=device)
        self.lr_t = torch.tensor([float(g["lr"]) for g in torch_optimizer.param_groups], dtype=torch.float32, device=device)
        self._lr_host = [float(g["lr"]) for g in torch_optimizer.param_groups]
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=device)
        self.partial = torch.empty(_lib.load().detr_sumsq_grid(self.total), dtype=torch.float32, device=device)
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)

    def sync_lr(self) -> None:
        """Upload the groups' current learning rates if an LR scheduler changed them (the kernels read them from device
        memory, so a captured graph needs this call -- not a re-capture -- to follow the schedule)."""
        cur = [float(g["lr"]) for g in self.opt.param_groups]
        if cur != self._lr_host:
            self.lr_t.copy_(torch.tensor(cur, dtype=torch.float32), non_blocking=True)
            self._lr_host = cur

    def step(self, max_norm: float, grad_div: float = 1.0) -> None:
        """One update from the gradients in `flat_g` (scaled by grad_div, e.g. 1 / world size, then clipped to max_norm)."""
        from . import _lib
        st = _lib.stream_ptr()
        # the sumsq launch also advances the step counter -- unless the gradient norm is not finite, in which case the AdamW
        # launches below leave parameters and moments untouched (a faulted batch is skipped, not applied)
        _lib.call("detr_sumsq_f32", self.flat_g.data_ptr(), self.total, self.partial.data_ptr(), self.sumsq.data_ptr(),
                  self.counter.data_ptr(), self.step_t.data_ptr(), st)
        for gi, (g, (start, n)) in enumerate(zip(self.opt.param_groups, self.ranges)):
            if n == 0:
                continue
            b1, b2 = g["betas"]
            off = start * 4
            _lib.call("detr_adamw_clip_f32", self.flat_p.data_ptr() + off, self.flat_g.data_ptr() + off, self.flat_m.data_ptr() + off,
                      self.flat_v.data_ptr() + off, n, self.lr_t.data_ptr() + 4 * gi, float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]),
                      self.step_t.data_ptr(), self.sumsq.data_ptr(), float(max_norm), float(grad_div), st)


class GraphedTrainStep:
    """The reference's step body (detr/train.py:258-267) captured ONCE in CUDA graphs and replayed: the eager step is
    host-launch-bound (~2 500 launches), the replay is GPU-bound.

    graph A: autocast forward, matcher + criterion, backward (+ gradient flattening when world > 1)
    [world > 1: one NCCL all-reduce of the flat gradient over NVLink -- the only collective besides num_boxes]
    graph B: (un-flatten, average) clip_grad_norm 1.0, fused AdamW step

    Inputs live in static buffers: `load(batch)` copies a host batch in (pinned, async) and repacks the ground truth
    into `StaticTargets`; attention dropout draws fresh masks at every replay through a device-side step counter.
    Requirements: fixed image size and batch size (one graph per shape), at most `gt_cap` boxes per image.
    """

    def __init__(self, model: nn.Module, criterion: nn.Module, optimizer, example: Dict, gt_cap: int = 100,
                 autocast_dtype=torch.bfloat16, max_grad_norm: float = 1.0, warmup: int = 3, flat_optimizer: bool = True,
                 restore_state: bool = True, check_faults: bool = True, accumulate: int = 1, overlap_allreduce: bool = False,
                 bucket_mb: float = 32.0):
        from . import attention
        from .targets import StaticTargets
        import torch.distributed as dist
        self.model, self.criterion, self.opt = model, criterion, optimizer
        self.dev = next(model.parameters()).device
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.autocast_dtype, self.max_grad_norm = autocast_dtype, max_grad_norm
        # gradient accumulation (detr/train.py:116,258 `accelerator.accumulate`): `accumulate` micro-batches per optimizer step; the
        # forward/backward graph then ADDS into the flat gradient buffer, the all-reduce and the update run once per step
        # (DDP's no_sync on the non-boundary micro-steps), and the 1 / accumulate of `accelerator.backward` rides in grad_div
        self.accumulate = max(int(accumulate), 1)
        self._micro = 0
        B = example["image"].shape[0]
        self.images = torch.empty_like(example["image"], device=self.dev, memory_format=torch.channels_last)
        self.heights = torch.empty(B, dtype=torch.int32, device=self.dev)
        self.widths = torch.empty(B, dtype=torch.int32, device=self.dev)
        self.targets = StaticTargets(B, model.config.num_object_queries, gt_cap, self.dev)
        self.step_counter = torch.zeros(1, dtype=torch.int64, device=self.dev)
        attention.set_dropout_step_tensor(self.step_counter)
        self.loss = torch.zeros((), device=self.dev)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.flat, self.flat_views = None, None
        # default: clip + AdamW as two HBM-speed launches on flat buffers (`optimizer` supplies groups and hyper-parameters)
        self.fopt = None
        if flat_optimizer and isinstance(optimizer, torch.optim.AdamW) and all(not g.get("amsgrad", False) for g in optimizer.param_groups):
            self.fopt = FlatAdamW(optimizer, self.dev)
            self.params = self.fopt.params
            self.flat, self.flat_views = self.fopt.flat_g, self.fopt.grad_views
        if self.accumulate > 1 and self.fopt is None:
            raise ValueError("gradient accumulation in GraphedTrainStep needs the flat optimizer (torch.optim.AdamW groups)")
        # Bucketed gradient all-reduce INSIDE the forward/backward graph, overlapped with the rest of backward (what DDP does for
        # the reference, detr/train.py:218,263): the flat gradient buffer is cut into contiguous buckets; when the last gradient
        # of a bucket has been produced (post-accumulate hooks: the transformer's bucket is complete ~5 ms before the ResNet
        # stem's), the bucket is copied into the flat buffer and its NCCL all-reduce starts on the communication stream while
        # autograd keeps going.  Captured in graph A as a fork / join.  OPT-IN: measured on 2 B200s it does not pay (12.83 ms vs
        # 12.74 ms for the single all-reduce between the two graphs -- the NCCL kernels take SMs from the backward kernels they
        # overlap with, and the whole 166 MB all-reduce is only ~0.3 ms over NVLink 5), and NCCL's asynchronous error handling has to
        # be switched off for the capture (TORCH_NCCL_ASYNC_ERROR_HANDLING=0).
        self._buckets = None
        if overlap_allreduce and self.world > 1 and self.fopt is not None and self.accumulate == 1:
            self._make_buckets(bucket_mb)
        self.load(example)
        # The warm-up iterations below are REAL optimizer steps (they have to be: allocator, cuDNN autotune, lazy attributes):
        # snapshot parameters, buffers and optimizer state first and put them back after the capture, so that training starts
        # from exactly the state the caller handed in (loss-curve parity with detr/train.py from init or from a checkpoint).
        snap = self._snapshot_state() if restore_state else None
        # warm-up on a side stream, then capture
        # warm-up and capture run on ONE high-priority stream: the captured kernels of the critical path (forward, input gradients)
        # then win SMs over the weight-gradient branch on the second stream (default priority) whenever both have blocks ready,
        # and autograd's AccumulateGrad nodes (created during warm-up) live on the stream the capture uses
        side = torch.cuda.Stream(priority=int(os.environ.get("DETR_B200_MAIN_PRIORITY", "-1")))
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._forward_backward()
                self._allreduce()
                self._update()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.opt.zero_grad(set_to_none=True)
        from . import _lib
        l0 = _lib.launch_count
        self.graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_a, stream=side):
            self._forward_backward()
        self.graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_b, pool=self.graph_a.pool(), stream=side):
            self._update()
        self.own_launches_per_step = _lib.launch_count - l0   # kernels of libdetr_b200.so inside one replay of both graphs
        if snap is not None:
            self._restore_state(snap)
        # device fault word of the matcher / criterion, read back asynchronously after every step and checked one step late
        self.check_faults = check_faults
        self._fault_host = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(2)]
        self._fault_event = [None, None]
        self._fault_i = 0
        self.skipped_steps = 0

    def _snapshot_state(self):
        with torch.no_grad():
            snap = {"params": self.fopt.flat_p.clone() if self.fopt is not None else [q.detach().clone() for q in self.params],
                    "buffers": [(b, b.detach().clone()) for b in self.model.buffers()],
                    "opt": {id(v): v.detach().clone() for st in self.opt.state.values() for v in st.values() if torch.is_tensor(v)}}
            if self.fopt is not None:
                snap["flat"] = (self.fopt.flat_m.clone(), self.fopt.flat_v.clone(), self.fopt.step_t.clone())
        return snap

    def _restore_state(self, snap) -> None:
        """In place (the captured graphs hold the addresses): parameters, buffers, moments, step counts, dropout step."""
        with torch.no_grad():
            if self.fopt is not None:
                self.fopt.flat_p.copy_(snap["params"])
                m, v, t = snap["flat"]
                self.fopt.flat_m.copy_(m); self.fopt.flat_v.copy_(v); self.fopt.step_t.copy_(t)
            else:
                for q, q0 in zip(self.params, snap["params"]):
                    q.copy_(q0)
            for b, b0 in snap["buffers"]:
                b.copy_(b0)
            for st in self.opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.copy_(snap["opt"][id(v)]) if id(v) in snap["opt"] else v.zero_()   # state created by the warm-up: back to zero
            self.step_counter.zero_()
            self.loss.zero_()
        torch.cuda.synchronize()

    # -- pieces -------------------------------------------------------------------------------------------
    def _make_buckets(self, bucket_mb: float) -> None:
        """Contiguous ranges of the flat gradient buffer of ~bucket_mb each, in flat (= parameter-group) order."""
        lim = int(bucket_mb * (1 << 20) / 4)
        self._buckets, cur = [], None
        offs = {}
        for v in self.flat_views:
            offs[id(v)] = v.storage_offset()
        for q, v in zip(self.params, self.flat_views):
            o = v.storage_offset()
            if cur is None or (o + q.numel() - cur["start"]) > lim:
                cur = {"start": o, "end": o, "params": [], "views": [], "pending": 0}
                self._buckets.append(cur)
            cur["params"].append(q); cur["views"].append(v); cur["end"] = max(cur["end"], o + q.numel())
        # a bucket's range must not overlap the next one's start (parameter groups are padded to 16 bytes: ranges are disjoint)
        self._works = []
        for b in self._buckets:
            for q in b["params"]:
                q.register_post_accumulate_grad_hook(lambda p_, b=b: self._grad_ready(b))

    def _grad_ready(self, b) -> None:
        if not self._hooks_live:
            return
        b["pending"] -= 1
        if b["pending"] == 0:
            import torch.distributed as dist
            torch._foreach_copy_(b["views"], [q.grad for q in b["params"]])
            self._works.append(dist.all_reduce(self.flat[b["start"]:b["end"]], async_op=True))

    _hooks_live = False

    def _forward_backward(self):
        for q in self.params:          # fresh gradient tensors (no accumulation kernels); Python-only, nothing is launched
            q.grad = None
        if self._buckets is not None:
            for b in self._buckets:
                b["pending"] = len(b["params"])
            self._works = []
            self._hooks_live = True
        self.step_counter.add_(1)
        with torch.autocast(device_type="cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            out = self.model(self.images, self.heights, self.widths)
        out = {k: v.float() for k, v in out.items()}
        losses = self.criterion(out, {"packed": self.targets.packed, "num_boxes": self.targets.num_boxes})
        loss = sum(v for k, v in losses.items() if k.startswith("loss"))
        # weight-gradient GEMMs on a parallel branch (gemm._SideStream): safe here because nothing reads a weight gradient
        # before backward() returns -- unless the bucketed all-reduce hooks are live
        prev = gemm.wgrad_side_stream(self._buckets is None and os.environ.get("DETR_B200_WGRAD_STREAM", "1") != "0")
        try:
            loss.backward()
        finally:
            gemm.wgrad_side_stream(prev)
        self.loss.copy_(loss.detach())
        if self._buckets is not None:
            self._hooks_live = False
            missing = [b for b in self._buckets if b["pending"] != 0]
            if missing:       # parameters that received no gradient in this graph (unused): reduce their buckets now, zeros included
                import torch.distributed as dist
                for b in missing:
                    torch._foreach_copy_(b["views"], [q.grad if q.grad is not None else torch.zeros_like(q) for q in b["params"]])
                    self._works.append(dist.all_reduce(self.flat[b["start"]:b["end"]], async_op=True))
            for w in self._works:   # join: the optimizer graph starts after the last bucket
                w.wait()
            return
        if self.world > 1 or self.fopt is not None:
            # one multi-tensor copy into the flat gradient / all-reduce buffer (torch.cat over ~330 gradients costs 0.8 ms, measured)
            if self.flat is None:
                self.flat = torch.empty(sum(q.numel() for q in self.params), dtype=torch.float32, device=self.dev)
                self.flat_views, o = [], 0
                for q in self.params:
                    # the view must look like the parameter (channels_last conv weights included) for the fused optimizer
                    self.flat_views.append(torch.as_strided(self.flat, q.shape, q.stride(), o))
                    o += q.numel()
            grads = [q.grad if q.grad is not None else torch.zeros_like(q) for q in self.params]
            if self.accumulate > 1:
                torch._foreach_add_(self.flat_views, grads)      # the buffer is zeroed at the start of every optimizer step
            else:
                torch._foreach_copy_(self.flat_views, grads)

    def _allreduce(self):
        if self.world > 1 and self._buckets is None:
            import torch.distributed as dist
            dist.all_reduce(self.flat)

    def _update(self):
        if self.fopt is not None:
            self.fopt.step(self.max_grad_norm, 1.0 / (self.world * self.accumulate))   # the data-parallel / accumulation mean rides in the clip coefficient
            return
        if self.world > 1:
            # the reduced gradients are consumed in place: .grad becomes a view of the flat buffer, averaged by ONE kernel
            self.flat.div_(float(self.world * self.accumulate))
            for q, v in zip(self.params, self.flat_views):
                q.grad = v
        torch.nn.utils.clip_grad_norm_(self.params, self.max_grad_norm)
        self.opt.step()

    # -- public -------------------------------------------------------------------------------------------
    def load(self, batch: Dict) -> int:
        """Copy a (host, ideally pinned) batch into the static buffers. Returns the bytes moved host->device."""
        import torch.distributed as dist
        self.images.copy_(batch["image"], non_blocking=True)
        self.heights.copy_(batch["height"], non_blocking=True)
        self.widths.copy_(batch["width"], non_blocking=True)
        n = self.targets.update(batch["class_idx"], batch["boxes_normalized"])
        self.targets.set_num_boxes(float(n))
        if self.world > 1 and getattr(self.criterion, "sync_num_boxes", True):
            t = self.targets.num_boxes.clone()
            dist.all_reduce(t)                         # the num_boxes all-reduce: 1 scalar, no host sync
            self.targets.num_boxes.copy_((t / self.world).clamp_(min=1.0))
        return (batch["image"].numel() * batch["image"].element_size() + 8 * self.heights.numel() + self.targets.bytes_per_update())

    def reset_optimizer_state(self) -> None:
        """Zero the moments and the step count (the constructor's warm-up iterations have advanced them)."""
        if self.fopt is not None:
            self.fopt.flat_m.zero_(); self.fopt.flat_v.zero_(); self.fopt.step_t.zero_()
        for st in self.opt.state.values():
            for v in st.values():
                if torch.is_tensor(v):
                    v.zero_()

    # double-buffered input pipeline: the H2D copy of batch i+1 runs on a copy stream under the compute of step i
    _copy_stream = None

    def prefetch(self, batch: Dict) -> None:
        """Start copying a (pinned) host batch's images into a staging buffer on a side stream; `commit()` makes it the
        current input.  The previous staging contents must have been committed."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._stage = torch.empty_like(self.images)
            self._ready, self._consumed = torch.cuda.Event(), torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream())
        self._copy_stream.wait_event(self._consumed)          # the last commit's device copy has read the staging buffer
        with torch.cuda.stream(self._copy_stream):
            self._stage.copy_(batch["image"], non_blocking=True)
            self._ready.record(self._copy_stream)
        self._pending = batch

    def commit(self) -> int:
        """Make the prefetched batch the input of the next `step()`: wait for its H2D copy, move it into the graph's static
        image buffer (device-to-device), upload sizes and repacked ground truth.  Returns the bytes moved host->device."""
        import torch.distributed as dist
        batch = self._pending
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ready)
        self.images.copy_(self._stage, non_blocking=True)
        self._consumed.record(cur)
        self.heights.copy_(batch["height"], non_blocking=True)
        self.widths.copy_(batch["width"], non_blocking=True)
        n = self.targets.update(batch["class_idx"], batch["boxes_normalized"])
        self.targets.set_num_boxes(float(n))
        if self.world > 1 and getattr(self.criterion, "sync_num_boxes", True):
            t = self.targets.num_boxes.clone()
            dist.all_reduce(t)
            self.targets.num_boxes.copy_((t / self.world).clamp_(min=1.0))
        return (batch["image"].numel() * batch["image"].element_size() + 8 * self.heights.numel() + self.targets.bytes_per_update())

    def _status_tensor(self):
        m = getattr(self.criterion, "matcher", None)
        return m.status_tensor(self.dev) if hasattr(m, "status_tensor") else None

    def poll_faults(self, wait: bool = False) -> None:
        """Raise the reference's exception (detr/utils.py:87-88 AssertionError, SciPy's ValueError) for a data fault recorded
        by an EARLIER step whose status read-back has completed (`wait=True`: block for the latest one).  The faulty step's
        update was skipped on the device (non-finite gradient norm) and the status word has been cleared, so training can
        continue past the exception if the caller chooses to."""
        from .matcher import raise_for_status
        for i in (0, 1):
            ev = self._fault_event[i]
            if ev is not None and (wait or ev.query()):
                ev.synchronize()
                self._fault_event[i] = None
                bits = int(self._fault_host[i][0])
                if bits:
                    self.skipped_steps += 1
                    raise_for_status(bits)

    def step(self) -> torch.Tensor:
        """One optimizer step on whatever `load()` put in the static buffers. Returns the (device) loss scalar.
        A data fault (degenerate box, NaN cost, bad label) poisons that step's losses with NaN; the optimizer kernels then
        skip the update, the status word is read back asynchronously, cleared, and the fault is raised by the NEXT
        `step()` / `poll_faults()` call -- no host synchronisation on the hot path, no sticky poisoning of later steps."""
        if self.check_faults:
            self.poll_faults()
        if self.accumulate > 1:
            if self._micro == 0:
                self.flat.zero_()
            self.graph_a.replay()
            self._micro += 1
            if self._micro < self.accumulate:
                return self.loss            # non-boundary micro-step: no collective, no update
            self._micro = 0
        else:
            self.graph_a.replay()
        st = self._status_tensor() if self.check_faults else None
        if st is not None:
            i = self._fault_i
            if self._fault_event[i] is not None:          # slot still in flight (two steps ago): it has long completed
                self.poll_faults(wait=True)
            self._fault_host[i].copy_(st, non_blocking=True)
            st.zero_()
            self._fault_event[i] = torch.cuda.Event()
            self._fault_event[i].record(torch.cuda.current_stream())
            self._fault_i = 1 - i
        self._allreduce()
        if self.fopt is not None:
            self.fopt.sync_lr()
        self.graph_b.replay()
        return self.loss
