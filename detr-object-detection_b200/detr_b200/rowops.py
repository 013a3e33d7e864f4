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
    _lib.call("detr_colsum_bf16", g.data_ptr(), g.stride(0), M, N, partial.data_ptr(), out.data_ptr(), _lib.stream_ptr())
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
            dw = (g2.t() @ x2).to(ctx.w_dtype)
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
