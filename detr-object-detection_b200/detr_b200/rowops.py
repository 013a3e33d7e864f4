"""Row-wise operations of the transformer blocks on libdetr_b200 kernels."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib


def colsum(g: torch.Tensor) -> torch.Tensor:
    """Column sums of a 2-D bf16 matrix -> fp32 (N,).  (Bias gradient of nn.Linear.)"""
    _lib.require_cuda(g, "colsum")
    M, N = g.shape
    if g.dtype != torch.bfloat16 or g.stride(1) != 1 or g.stride(0) % 8 or g.data_ptr() % 16 or N % 8:
        return g.float().sum(0)  # outside the kernel's contract (e.g. the 92-wide class head: 92 % 8 != 0)
    chunks = _lib.load().detr_colsum_chunks(M, N)
    partial = torch.empty(chunks * N, dtype=torch.float32, device=g.device)
    out = torch.empty(N, dtype=torch.float32, device=g.device)
    _lib.call("detr_colsum_bf16", g.data_ptr(), g.stride(0), M, N, partial.data_ptr(), out.data_ptr(),
              _lib.zero_counters(g.device).data_ptr(), _lib.stream_ptr())
    return out


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b with cuBLASLt (bias in the GEMM epilogue) and a backward whose bias gradient is ONE pass of the
    colsum kernel instead of ATen's generic reduction."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        # x bf16; weight / bias are the fp32 parameters: cast here so that their gradients come back in fp32
        w16 = weight.to(x.dtype)
        ctx.save_for_backward(x, w16)
        ctx.has_bias = bias is not None
        ctx.w_dtype = weight.dtype
        return F.linear(x, w16, None if bias is None else bias.to(x.dtype))

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        x2 = x.reshape(-1, x.shape[-1])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = (g2 @ weight).view(x.shape)
        if ctx.needs_input_grad[1]:
            dw = torch.mm(g2.t(), x2, out_dtype=ctx.w_dtype) if ctx.w_dtype == torch.float32 else (g2.t() @ x2).to(ctx.w_dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(g2 if g2.is_contiguous() or g2.stride(1) == 1 else g2.contiguous())
        return dx, dw, db


def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None) -> torch.Tensor:
    """nn.functional.linear under bf16 autocast with the fast bias-gradient; plain F.linear otherwise."""
    if torch.is_autocast_enabled() and x.is_cuda:
        dt = torch.get_autocast_dtype("cuda")
        if dt == torch.bfloat16:
            with torch.autocast("cuda", enabled=False):
                return _LinearFn.apply(x.to(dt), weight, bias)
    return F.linear(x, weight, bias)


# ------------------------------------------------------------------------------------------------ fused LayerNorm
_DT = {torch.float32: 0, torch.bfloat16: 1}


class _LayerNormAdd(torch.autograd.Function):
    """(y, y2) = (LN(x), LN(x) + addend) in one kernel; y / y2 in `out_dtype`.  `addend` is fp32 (B, R, C) (any batch /
    row strides, e.g. an expanded (R, C) embedding) or None; `want_y=False` skips writing y (decoder cross-attention
    only needs the query = LN(x) + query_embed)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, addend, eps, out_dtype, want_y):
        shape = x.shape
        C = shape[-1]
        x2 = x.reshape(-1, C)
        if x2.stride(1) != 1 or x2.stride(0) % 8 or x2.data_ptr() % 16:
            x2 = x2.contiguous()
        rows = x2.shape[0]
        dev = x.device
        y = torch.empty(shape, dtype=out_dtype, device=dev) if want_y else None
        y2 = torch.empty(shape, dtype=out_dtype, device=dev) if addend is not None else None
        stats = torch.empty(2, rows, dtype=torch.float32, device=dev)
        a_ptr, a_sb, a_sr, rpb = None, 0, 0, rows
        if addend is not None:
            a = addend if addend.dtype == torch.float32 else addend.float()
            if a.dim() != 3 or a.shape != shape or a.stride(2) != 1:
                a = a.expand(shape).contiguous() if a.shape != shape else a.contiguous()
            a_ptr, a_sb, a_sr, rpb = a.data_ptr(), a.stride(0), a.stride(1), shape[1]
            ctx.addend_ref = a  # keep alive until the launch is enqueued (same stream: safe to drop afterwards)
        g32, b32 = gamma.float(), beta.float()
        _lib.call("detr_layernorm_fwd", x2.data_ptr(), _DT[x2.dtype], x2.stride(0), g32.data_ptr(), b32.data_ptr(),
                  a_ptr, 0, a_sb, a_sr, rpb, _lib.ptr(y), _lib.ptr(y2), _DT[out_dtype], stats[0].data_ptr(), stats[1].data_ptr(),
                  rows, C, float(eps), _lib.stream_ptr())
        ctx.save_for_backward(x2, g32, stats)
        ctx.shape, ctx.has_addend, ctx.want_y = shape, addend is not None, want_y
        ctx.addend_dtype = addend.dtype if addend is not None else None
        ctx.x_dtype = x.dtype
        ctx.mark_non_differentiable()
        return y, y2

    @staticmethod
    def backward(ctx, dy, dy2):
        x2, g32, stats = ctx.saved_tensors
        rows, C = x2.shape
        gs = [t for t in (dy, dy2) if t is not None]
        if not gs:
            return None, None, None, None, None, None, None
        gdt = torch.float32 if any(t.dtype == torch.float32 for t in gs) else torch.bfloat16
        prep = lambda t: None if t is None else t.to(gdt).reshape(rows, C).contiguous()
        dyc, dy2c = prep(dy), prep(dy2)
        dx = torch.empty(rows, C, dtype=x2.dtype, device=x2.device)
        grid = _lib.load().detr_layernorm_grid(rows)
        partial = torch.empty(grid * 2 * C, dtype=torch.float32, device=x2.device)
        dgb = torch.empty(2, C, dtype=torch.float32, device=x2.device)
        _lib.call("detr_layernorm_bwd", _lib.ptr(dyc), _lib.ptr(dy2c), _DT[gdt], x2.data_ptr(), _DT[x2.dtype], x2.stride(0),
                  g32.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), dx.data_ptr(), partial.data_ptr(),
                  dgb[0].data_ptr(), dgb[1].data_ptr(), _lib.zero_counters(x2.device).data_ptr(), rows, C, _lib.stream_ptr())
        d_add = dy2.to(ctx.addend_dtype) if (ctx.has_addend and dy2 is not None and ctx.needs_input_grad[3]) else None
        return dx.view(ctx.shape), dgb[0], dgb[1], d_add, None, None, None


def layer_norm_add(x, norm: torch.nn.LayerNorm, addend=None, want_y: bool = True):
    """-> (LN(x), LN(x) + addend).  bf16 outputs under bf16 autocast (what the following GEMMs consume), else x.dtype."""
    out_dtype = torch.bfloat16 if (torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16) else x.dtype
    if x.dtype not in _DT or out_dtype not in _DT:
        raise TypeError(f"layer_norm_add supports float32 / bfloat16, got {x.dtype}")
    _lib.require_cuda(x, "layer_norm_add")
    with torch.autocast("cuda", enabled=False):
        return _LayerNormAdd.apply(x, norm.weight, norm.bias, addend, norm.eps, out_dtype, want_y)
