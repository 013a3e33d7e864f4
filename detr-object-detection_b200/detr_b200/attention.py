"""Attention core of ScaledDotProductAttention (detr/model.py:317-352) on tcgen05/TMA kernels.

q (B,L,C), k/v (B,S,C) are the *projected* bf16 tensors in nn.Linear's layout; heads are 32-channel slices, so
no view/transpose/contiguous round trip is needed.  Returns (B,L,C) bf16.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

HEAD_DIM = 32


def _tma_ok(t: torch.Tensor) -> torch.Tensor:
    """TMA needs channel stride 1, 16-byte aligned base and row/batch strides that are multiples of 8 elements."""
    if t.dtype != torch.bfloat16:
        t = t.to(torch.bfloat16)
    if t.stride(2) != 1 or t.data_ptr() % 16 or t.stride(0) % 8 or t.stride(1) % 8:
        t = t.contiguous()
    return t


def _mask_bytes(m: Optional[torch.Tensor], shape, name: str) -> Optional[torch.Tensor]:
    if m is None:
        return None
    if tuple(m.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(m.shape)}")
    m = m.to(torch.bool) if m.dtype != torch.bool else m
    return m.contiguous().view(torch.uint8)


def attention_forward(q, k, v, key_padding_mask=None, attention_mask=None, dropout_p: float = 0.0, seed: int = 0,
                      seed_tensor: Optional[torch.Tensor] = None, residual: bool = False):
    """Raw forward launch -> (out bf16 (B,L,C), lse fp32 (B,nh,L)).  With `residual` the output is (2, B, L, C): out[0] is
    the attention output, out[1] its bf16 rounding residual (what `attention_backward(out_lo=)` uses for an exact delta)."""
    _lib.require_cuda(q, "attention")
    B, L, C = q.shape
    S = k.shape[1]
    if C % HEAD_DIM:
        raise ValueError(f"hidden size {C} is not a multiple of the head size {HEAD_DIM} this kernel is built for")
    nh = C // HEAD_DIM
    q, k, v = _tma_ok(q), _tma_ok(k), _tma_ok(v)
    kpm = _mask_bytes(key_padding_mask, (B, S), "key_padding_mask")
    am = _mask_bytes(attention_mask, (L, S), "attention_mask")
    out2 = torch.empty(2 if residual else 1, B, L, C, dtype=torch.bfloat16, device=q.device)
    out = out2[0]
    lse = torch.empty(B, nh, L, dtype=torch.float32, device=q.device)
    ws = torch.empty(_lib.load().detr_attention_fwd_workspace_floats(B, nh, L, S), dtype=torch.float32, device=q.device)
    _lib.call(
        "detr_attention_fwd_bf16",
        q.data_ptr(), q.stride(0), q.stride(1), k.data_ptr(), k.stride(0), k.stride(1), v.data_ptr(), v.stride(0), v.stride(1),
        out.data_ptr(), out.stride(0), out.stride(1), out2[1].data_ptr() if residual else None, lse.data_ptr(), ws.data_ptr(),
        _lib.ptr(kpm), kpm.stride(0) if kpm is not None else 0,
        _lib.ptr(am), B, nh, L, S, float(dropout_p), int(seed) & 0xFFFFFFFFFFFFFFFF, _lib.ptr(seed_tensor), _lib.stream_ptr(),
        tag=(B, nh, L, S), launches=2 if -(-L // 128) * nh * B > _lib.num_sms() else 1)   # + combine when items are split between CTAs
    return (out2 if residual else out), lse


def attention_backward(d_out, q, k, v, out, lse, key_padding_mask=None, attention_mask=None, dropout_p: float = 0.0,
                       seed: int = 0, seed_tensor: Optional[torch.Tensor] = None, outs: Optional[tuple] = None,
                       out_lo: Optional[torch.Tensor] = None):
    """Raw backward launches -> (dq, dk, dv) bf16 with the shapes of q, k, v.  `outs`: caller-provided (dq, dk, dv) bf16
    buffers (channel stride 1, 16-byte aligned, row / batch strides multiples of 8 -- e.g. the two halves of one (B, L, 2C)
    gradient of a fused q/k projection)."""
    B, L, C = q.shape
    S = k.shape[1]
    nh = C // HEAD_DIM
    q, k, v, out, d_out = _tma_ok(q), _tma_ok(k), _tma_ok(v), _tma_ok(out), _tma_ok(d_out)
    if out_lo is not None and (out_lo.dtype != torch.bfloat16 or out_lo.shape != out.shape or out_lo.stride() != out.stride() or out_lo.data_ptr() % 16):
        raise ValueError("attention_backward: out_lo must be a bf16 tensor with out's shape and strides")
    kpm = _mask_bytes(key_padding_mask, (B, S), "key_padding_mask")
    am = _mask_bytes(attention_mask, (L, S), "attention_mask")
    if outs is None:
        dq = torch.empty(B, L, C, dtype=torch.bfloat16, device=q.device)
        dkv = torch.empty(2, B, S, C, dtype=torch.bfloat16, device=q.device)
        dk, dv = dkv[0], dkv[1]
    else:
        dq, dk, dv = outs
        for t, n in ((dq, L), (dk, S), (dv, S)):
            if (t.dtype != torch.bfloat16 or tuple(t.shape) != (B, n, C) or t.stride(2) != 1 or t.data_ptr() % 16
                    or t.stride(0) % 8 or t.stride(1) % 8):
                raise ValueError("attention_backward: `outs` must be bf16 (B, rows, C) buffers with channel stride 1, 16-byte aligned")
    delta = torch.empty(B, nh, L, dtype=torch.float32, device=q.device)
    dq_part = torch.empty(_lib.load().detr_attention_bwd_workspace_floats(B, nh, L, S), dtype=torch.float32, device=q.device)
    st = lambda t: (t.data_ptr(), t.stride(0), t.stride(1))
    _lib.call(
        "detr_attention_bwd_bf16",
        *st(q), *st(k), *st(v), *st(out), _lib.ptr(out_lo), *st(d_out), lse.data_ptr(), delta.data_ptr(), dq_part.data_ptr(), *st(dq), *st(dk), *st(dv),
        _lib.ptr(kpm), kpm.stride(0) if kpm is not None else 0, _lib.ptr(am), B, nh, L, S, float(dropout_p),
        int(seed) & 0xFFFFFFFFFFFFFFFF, _lib.ptr(seed_tensor), _lib.stream_ptr(), tag=(B, nh, L, S),
        launches=4 if -(-S // 128) * nh * B > _lib.num_sms() else 3)   # delta, main, dQ reduction (+ dK/dV reduction when items are split)
    return dq, dk, dv


class _FlashAttentionQK(torch.autograd.Function):
    """Self-attention on the output of ONE fused q/k projection: qk (B, L, 2C), q = qk[..., :C], k = qk[..., C:].  The kernels
    take the two halves as strided views in both directions: backward writes dq and dk straight into the halves of one
    (B, L, 2C) buffer.  (With q and k as autograd slices every layer paid two zero-fills, two slice copies and one add of
    (B, L, 2C) in backward.)"""

    @staticmethod
    def forward(ctx, qk, v, key_padding_mask, attention_mask, dropout_p, seed, seed_tensor):
        C = qk.shape[-1] // 2
        out2, lse = attention_forward(qk[..., :C], qk[..., C:], v, key_padding_mask, attention_mask, dropout_p, seed, seed_tensor, residual=True)
        ctx.save_for_backward(qk, v, out2, lse, key_padding_mask, attention_mask, seed_tensor)
        ctx.dropout_p, ctx.seed = dropout_p, seed
        return out2[0]

    @staticmethod
    def backward(ctx, d_out):
        qk, v, out2, lse, kpm, am, seed_tensor = ctx.saved_tensors
        B, L, C2 = qk.shape
        C = C2 // 2
        dqk = torch.empty(B, L, C2, dtype=torch.bfloat16, device=qk.device)
        dv = torch.empty(B, v.shape[1], C, dtype=torch.bfloat16, device=qk.device)
        attention_backward(d_out, qk[..., :C], qk[..., C:], v, out2[0], lse, kpm, am, ctx.dropout_p, ctx.seed, seed_tensor,
                           outs=(dqk[..., :C], dqk[..., C:], dv), out_lo=out2[1])
        return dqk.to(qk.dtype), dv.to(v.dtype), None, None, None, None, None


class _FlashAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, key_padding_mask, attention_mask, dropout_p, seed, seed_tensor):
        out2, lse = attention_forward(q, k, v, key_padding_mask, attention_mask, dropout_p, seed, seed_tensor, residual=True)
        ctx.save_for_backward(q, k, v, out2, lse, key_padding_mask, attention_mask, seed_tensor)
        ctx.dropout_p, ctx.seed = dropout_p, seed
        ctx.in_dtypes = (q.dtype, k.dtype, v.dtype)
        return out2[0]

    @staticmethod
    def backward(ctx, d_out):
        q, k, v, out2, lse, kpm, am, seed_tensor = ctx.saved_tensors
        dq, dk, dv = attention_backward(d_out, q, k, v, out2[0], lse, kpm, am, ctx.dropout_p, ctx.seed, seed_tensor, out_lo=out2[1])
        tq, tk, tv = ctx.in_dtypes
        return dq.to(tq), dk.to(tk), dv.to(tv), None, None, None, None, None


def flash_attention(q, k, v, key_padding_mask: Optional[torch.Tensor] = None,
                    attention_mask: Optional[torch.Tensor] = None, dropout_p: float = 0.0,
                    seed: Optional[int] = None) -> torch.Tensor:
    """Differentiable attention core: softmax(q k^T / sqrt(32) + masks) -> dropout -> @ v, heads = 32-channel slices.

    Dropout seed = host draw (follows torch.manual_seed, no device sync) + the device-side step counter registered with
    `set_dropout_step_tensor` (so a captured CUDA graph draws a fresh mask at every replay)."""
    if dropout_p > 0.0 and seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    st = _STEP_TENSOR if (dropout_p > 0.0 and _STEP_TENSOR is not None and _STEP_TENSOR.device == q.device) else None
    return _FlashAttention.apply(q, k, v, key_padding_mask, attention_mask, float(dropout_p), int(seed or 0), st)


def flash_attention_qk(qk, v, key_padding_mask: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                       dropout_p: float = 0.0, seed: Optional[int] = None) -> torch.Tensor:
    """`flash_attention(qk[..., :C], qk[..., C:], v, ...)` for the (B, L, 2C) output of a fused q/k projection, as one autograd
    node (see _FlashAttentionQK)."""
    if qk.dim() != 3 or qk.shape[-1] % 2 or v.shape[-1] * 2 != qk.shape[-1]:
        raise ValueError("flash_attention_qk: qk must be (B, L, 2C) and v (B, S, C)")
    if dropout_p > 0.0 and seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    st = _STEP_TENSOR if (dropout_p > 0.0 and _STEP_TENSOR is not None and _STEP_TENSOR.device == qk.device) else None
    return _FlashAttentionQK.apply(qk, v, key_padding_mask, attention_mask, float(dropout_p), int(seed or 0), st)


class _FlashAttentionQKV(torch.autograd.Function):
    """Self-attention on the output of ONE fused q|k|v projection: qkv (B, L, 3C).  The kernels take the three column blocks as
    strided views in both directions: backward writes dq, dk, dv straight into one (B, L, 3C) gradient buffer, which is what
    the projection's input-gradient / weight-gradient GEMMs consume."""

    @staticmethod
    def forward(ctx, qkv, key_padding_mask, attention_mask, dropout_p, seed, seed_tensor):
        C = qkv.shape[-1] // 3
        out2, lse = attention_forward(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], key_padding_mask, attention_mask, dropout_p, seed,
                                      seed_tensor, residual=True)
        ctx.save_for_backward(qkv, out2, lse, key_padding_mask, attention_mask, seed_tensor)
        ctx.dropout_p, ctx.seed = dropout_p, seed
        return out2[0]

    @staticmethod
    def backward(ctx, d_out):
        qkv, out2, lse, kpm, am, seed_tensor = ctx.saved_tensors
        B, L, C3 = qkv.shape
        C = C3 // 3
        dqkv = torch.empty(B, L, C3, dtype=torch.bfloat16, device=qkv.device)
        attention_backward(d_out, qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], out2[0], lse, kpm, am, ctx.dropout_p, ctx.seed, seed_tensor,
                           outs=(dqkv[..., :C], dqkv[..., C:2 * C], dqkv[..., 2 * C:]), out_lo=out2[1])
        return dqkv, None, None, None, None, None


def flash_attention_qkv(qkv, key_padding_mask: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                        dropout_p: float = 0.0, seed: Optional[int] = None) -> torch.Tensor:
    """Self-attention core on the (B, L, 3C) bf16 output of a fused q|k|v projection, one autograd node."""
    if qkv.dim() != 3 or qkv.shape[-1] % 3 or qkv.dtype != torch.bfloat16:
        raise ValueError("flash_attention_qkv: qkv must be a bf16 (B, L, 3C) tensor")
    if dropout_p > 0.0 and seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    st = _STEP_TENSOR if (dropout_p > 0.0 and _STEP_TENSOR is not None and _STEP_TENSOR.device == qkv.device) else None
    return _FlashAttentionQKV.apply(qkv, key_padding_mask, attention_mask, float(dropout_p), int(seed or 0), st)


_STEP_TENSOR: Optional[torch.Tensor] = None


def set_dropout_step_tensor(t: Optional[torch.Tensor]) -> None:
    """Register a device uint64/int64 scalar that is ADDED to every dropout seed inside the kernels.  A training step
    captured in a CUDA graph increments it once per replay; eager code can leave it unset."""
    global _STEP_TENSOR
    if t is not None and (t.numel() != 1 or t.dtype not in (torch.int64, torch.uint64) or not t.is_cuda):
        raise ValueError("dropout step tensor must be a CUDA int64/uint64 scalar")
    _STEP_TENSOR = t
