"""Pre-LN block pieces of the encoder / decoder layers (detr/model.py:165-183, 220-225) as autograd functions over the
tcgen05 GEMM kernels of csrc/gemm.cu (through `gemm.py`).  Under bf16 autocast a layer is

    qkv = LnProj(x, norm1, + pos / query embedding on the q|k columns)       1 launch   (detr/model.py:221-222,312-314)
    y   = flash attention(qkv)                                                1-2        (:317-352)
    x   = ProjRes(y, output_proj, residual = x)                               1          (:354-355,223)
    x   = LnFfn(x, norm2, ffn)                                                2          (:224,405-411)

with no cuBLAS / ATen kernel in between; the backward functions use the same kernels (input gradients: `gemm(b_kn=True)` on
the weight itself; weight + bias gradients: `gemm_wgrad`; GELU / dropout backward in the input-gradient GEMM's epilogue) plus
the LayerNorm backward and dropout-mask row kernels of csrc/rowops.cu.

Shapes: activations are (B, L, C) tensors handled as (B*L, C) matrices; weights are the bf16 shadows kept by
`rowops.ShadowedLinears` (refreshed once per forward), biases stay fp32."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, gemm as G
from .rowops import _DT, _ep_seed

C_MODEL = 256   # the LayerNorm-prologue kernel is built for hidden size 256 (8 heads of 32)


# ---- hand-over between a block's LayerNorm backward and the PREVIOUS block's tail backward ---------------------------------
# x_k = res + dropout(z) (a tail: _ProjRes / _LnFfn forward) is consumed by exactly one LayerNorm-prologue function of the next
# block.  In backward that function computes dx_k, which is precisely the gradient the tail receives -- and what the tail needs
# first is dz = bf16(mask(dx_k) / (1 - p)) plus its column sums (the bias gradient).  The LayerNorm backward kernel can write
# both in the same pass if it knows the tail's dropout parameters: tails register them under the data pointer of their output,
# the consuming forward picks them up, and the backward leaves (dz, dbias) under the data pointer of dx_k for the tail's
# backward to find.  Anything that does not match falls back to the separate row kernel.
_DZ_FOR: dict = {}      # data_ptr of dx -> (dx tensor, dz bf16 (M, C), dbias fp32 (C,))


def _register_tail(out: torch.Tensor, p: float, seed: int, seed_t) -> torch.Tensor:
    """Forward side of the hand-over: the tail's dropout parameters travel as a Python attribute of its output tensor (the very
    object the next block's wrapper function receives)."""
    out._detr_tail = (float(p), int(seed), seed_t)
    return out


def _claim_tail(x: torch.Tensor):
    return getattr(x, "_detr_tail", None)


def _take_dz(g: torch.Tensor):
    """(dz, dbias) left by the LayerNorm backward that produced exactly this gradient tensor, or None.  The entry keeps its dx
    alive, so while it exists no other tensor can have that data pointer: pointer + shape + dtype identify the tensor even when
    autograd hands over a fresh Python wrapper.  (A gradient that autograd had to accumulate from several consumers is a new
    tensor with a new pointer: no hit, the caller falls back to the row kernel.)"""
    hit = _DZ_FOR.pop(g.data_ptr(), None)
    if hit is not None and hit[0].shape == g.shape and hit[0].dtype == g.dtype and g.is_contiguous():
        return hit[1], hit[2]
    return None


def reset_handover() -> None:
    """Drop stale hand-over entries (a backward that was interrupted)."""
    _DZ_FOR.clear()


def _ln_backward(dy, dy2, dres, x2, gamma32, stats, tail=None):
    """LayerNorm backward on (rows, C) matrices: dx (x's dtype), dgamma, dbeta (fp32).  dy / dy2: gradients of the two
    normalised operands (either may be None), dres: gradient of the residual branch (added to dx in the same pass).
    `tail` = (p, seed, seed_t) of the block tail that produced x: the kernel then also writes that tail's masked bf16 gradient and
    bias gradient, left in `_DZ_FOR` for its backward."""
    rows, C = x2.shape
    gs = [t for t in (dy, dy2) if t is not None]
    gdt = torch.float32 if any(t.dtype == torch.float32 for t in gs) else torch.bfloat16
    prep = lambda t: None if t is None else (t if (t.dtype == gdt and t.is_contiguous()) else t.to(gdt).contiguous())
    dy, dy2 = prep(dy), prep(dy2)
    if dres is not None:
        dres = dres.reshape(rows, C)
        if dres.dtype != x2.dtype or not dres.is_contiguous():
            dres = dres.to(x2.dtype).contiguous()
    dx = torch.empty(rows, C, dtype=x2.dtype, device=x2.device)
    grid = _lib.load().detr_layernorm_grid(rows)
    with_tail = tail is not None and C == C_MODEL
    NP = 3 if with_tail else 2
    partial = torch.empty(grid * NP * C, dtype=torch.float32, device=x2.device)
    dgb = torch.empty(NP, C, dtype=torch.float32, device=x2.device)
    # the fold of the per-CTA partials into dgamma / dbeta (/ dbias) feeds parameter gradients only: inside a backward pass with the
    # second stream enabled (gemm._SideStream) it leaves the critical path, otherwise the same C call launches it
    side = G._SIDE.fork(x2.device, (partial,))      # outputs are never kept: see gemm._SideStream
    dg_ptr = dgb[0].data_ptr() if side is None else None
    if with_tail:
        p, seed, seed_t = tail
        dz = torch.empty(rows, C, dtype=torch.bfloat16, device=x2.device)
        _lib.call("detr_layernorm_bwd_tail", _lib.ptr(dy), _lib.ptr(dy2), _DT[gdt], _lib.ptr(dres), x2.data_ptr(), _DT[x2.dtype], x2.stride(0),
                  gamma32.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), dx.data_ptr(), partial.data_ptr(),
                  dg_ptr, dgb[1].data_ptr(), _lib.zero_counters(x2.device).data_ptr(), rows, C, dz.data_ptr(), dgb[2].data_ptr(),
                  float(p), seed & 0xFFFFFFFFFFFFFFFF, _lib.ptr(seed_t), _lib.stream_ptr(), launches=2 if side is None else 1)
        if len(_DZ_FOR) > 64:          # entries whose consumer never came (a branch of the graph that was not a block tail)
            _DZ_FOR.clear()
        _DZ_FOR[dx.data_ptr()] = (dx, dz, dgb[2])
    else:
        _lib.call("detr_layernorm_bwd", _lib.ptr(dy), _lib.ptr(dy2), _DT[gdt], _lib.ptr(dres), x2.data_ptr(), _DT[x2.dtype], x2.stride(0),
                  gamma32.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), dx.data_ptr(), partial.data_ptr(),
                  dg_ptr, dgb[1].data_ptr(), _lib.zero_counters(x2.device).data_ptr(), rows, C, _lib.stream_ptr(),
                  launches=2 if side is None else 1)
    if side is not None:
        G._SIDE.refork(x2.device)      # the fold waits for the row pass just launched
        with torch.cuda.stream(side):
            _lib.call("detr_layernorm_bwd_fold", partial.data_ptr(), rows, C, dgb[0].data_ptr(), dgb[1].data_ptr(),
                      dgb[2].data_ptr() if with_tail else None, _lib.stream_ptr())
        G._SIDE.debug_after_launch(x2.device)
    return dx, dgb[0], dgb[1]


def _drop_mask_bf16(g2: torch.Tensor, p: float, seed: int, seed_t, want_db: bool):
    """dy (bf16) = dropout_mask(g) / (1 - p) for a block tail's incoming gradient g (M, N), and the bias gradient
    db = column sums of dy when asked.  With p == 0 and a bf16 gradient nothing is launched."""
    M, N = g2.shape
    if p <= 0.0 and g2.dtype == torch.bfloat16 and g2.is_contiguous():
        return g2, None
    if g2.dtype not in _DT:
        g2 = g2.float()
    if not g2.is_contiguous():
        g2 = g2.contiguous()
    dy = torch.empty(M, N, dtype=torch.bfloat16, device=g2.device)
    chunks = _lib.load().detr_epilogue_chunks(M, N)
    partial = torch.empty(chunks * N, dtype=torch.float32, device=g2.device)
    db = torch.empty(N, dtype=torch.float32, device=g2.device)
    _lib.call("detr_epilogue_bwd", 0, g2.data_ptr(), _DT[g2.dtype], None, dy.data_ptr(), partial.data_ptr(), db.data_ptr(),
              _lib.zero_counters(g2.device).data_ptr(), M, N, float(p), seed & 0xFFFFFFFFFFFFFFFF, _lib.ptr(seed_t), _lib.stream_ptr())
    return dy, (db if want_db else None)


def _split_rows(t: Optional[torch.Tensor], sizes):
    if t is None:
        return [None] * len(sizes)
    return list(torch.split(t, sizes, 0))


def _as_grad(dx: torch.Tensor, shape):
    """dx (rows, C) as the (B, L, C) gradient handed to autograd; a pending hand-over entry follows the tensor object that the
    previous block's tail will actually receive."""
    out = dx.view(shape)
    hit = _DZ_FOR.get(dx.data_ptr())
    if hit is not None:
        _DZ_FOR[dx.data_ptr()] = (out, hit[1], hit[2])
    return out


class _LnProj(torch.autograd.Function):
    """out (B, L, N) bf16 = (LN(x) [+ addend on the first n_pos_end columns' operand]) @ W^T + b, W = the stacked weights.
    Second output: x itself (used as the block's residual input; its gradient comes back as `dres` and is added inside the
    LayerNorm backward kernel instead of by an autograd accumulation kernel)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, addend, eps, n_pos_end, w16, b32, tail, n_w, *params):
        weights, biases = params[:n_w], params[n_w:]
        B, L, C = x.shape
        x2 = x.reshape(B * L, C)
        if x2.stride(1) != 1 or x2.stride(0) % 8 or x2.data_ptr() % 16:
            x2 = x2.contiguous()
        if w16 is None:
            w16 = (weights[0] if n_w == 1 else torch.cat(weights, 0)).to(torch.bfloat16)
        if b32 is None:
            b32 = (biases[0] if n_w == 1 else torch.cat(biases, 0)).float()
        add_t, sb, sr = None, 0, 0
        if n_pos_end > 0:
            add_t = addend if addend.dtype == torch.float32 else addend.float()
            sb = add_t.stride(0) if (add_t.dim() == 3 and add_t.shape[0] > 1) else 0
            ok = (add_t.dim() == 3 and add_t.shape[1:] == (L, C) and add_t.stride(2) == 1 and add_t.stride(1) % 4 == 0 and add_t.data_ptr() % 16 == 0
                  and (B == 1 or sb == L * add_t.stride(1) or (sb == 0 and L >= 8)))     # dense block or batch broadcast
            if not ok:
                add_t = add_t.expand(B, L, C).contiguous()
                sb = add_t.stride(0)
            sr = add_t.stride(1)
        g32, be32 = gamma.float(), beta.float()
        out, a_plain, a_pos, stats = G.gemm_ln(x2, g32, be32, eps, w16, addend=add_t, rows_per_batch=L, add_sb=sb, add_sr=sr,
                                               n_pos_end=n_pos_end, bias=b32)
        ctx.save_for_backward(x2, g32, stats, a_plain, a_pos, w16)
        ctx.n_pos_end, ctx.n_w, ctx.shape = n_pos_end, n_w, (B, L, C)
        ctx.splits = [w.shape[0] for w in weights]
        ctx.addend_grad = addend is not None and n_pos_end > 0 and addend.requires_grad
        ctx.addend_dtype = addend.dtype if addend is not None else None
        ctx.tail = tail if x2.data_ptr() == x.data_ptr() else None
        ctx.mark_non_differentiable()
        return out.view(B, L, -1), x

    @staticmethod
    def backward(ctx, dout, dres):
        x2, g32, stats, a_plain, a_pos, w16 = ctx.saved_tensors
        B, L, C = ctx.shape
        N, npe = w16.shape[0], ctx.n_pos_end
        d2 = dout.reshape(B * L, N)
        if d2.dtype != torch.bfloat16:
            d2 = d2.to(torch.bfloat16)
        # weight / bias gradients: q|k rows contract with LN(x)+addend, the remaining rows with LN(x)
        if npe > 0 and npe < N:
            dw, db = G.gemm_wgrad(d2, a_pos, a_plain, n_switch=npe)
        else:
            dw, db = G.gemm_wgrad(d2, a_pos if npe > 0 else a_plain)
        # gradient of the normalised operand(s): one GEMM on the whole stacked weight unless the addend needs its own gradient
        d_add = None
        if ctx.addend_grad:
            # the addend's gradient IS g_pos: written by the GEMM in the addend's dtype (fp32 for the query embedding), so that
            # no cast kernel sits between this node and autograd's accumulation of the 12 uses of the embedding
            gdt = ctx.addend_dtype if ctx.addend_dtype in (torch.float32, torch.bfloat16) else torch.bfloat16
            g_pos = G.gemm(d2[:, :npe], w16[:npe], b_kn=True, out_dtype=gdt)
            g_plain = G.gemm(d2[:, npe:], w16[npe:], b_kn=True, out_dtype=gdt) if npe < N else None
            dx, dgam, dbet = _ln_backward(g_plain, g_pos, dres, x2, g32, stats, ctx.tail)
            d_add = g_pos.view(B, L, C).to(ctx.addend_dtype)
        else:
            g_all = G.gemm(d2, w16, b_kn=True)
            dx, dgam, dbet = _ln_backward(g_all, None, dres, x2, g32, stats, ctx.tail)
        return (_as_grad(dx, (B, L, C)), dgam, dbet, d_add, None, None, None, None, None, None,
                *_split_rows(dw, ctx.splits), *_split_rows(db, ctx.splits))


def ln_proj(x, norm: torch.nn.LayerNorm, linears, addend=None, n_pos_end: int = 0, w16=None, b32=None):
    """-> (proj (B, L, sum N_i) bf16, x).  `linears`: nn.Linear modules stacked along the output dimension; the first
    `n_pos_end` output columns are computed from LN(x) + addend."""
    with torch.autocast("cuda", enabled=False):
        return _LnProj.apply(x, norm.weight, norm.bias, addend, float(norm.eps), int(n_pos_end), w16, b32, _claim_tail(x), len(linears),
                             *[l.weight for l in linears], *[l.bias for l in linears])


class _ProjRes(torch.autograd.Function):
    """out = res + dropout(a @ W^T + b)  (attention output projection + residual, detr/model.py:354-355,223)."""

    @staticmethod
    def forward(ctx, a, res, w16, weight, bias, p, seed, seed_t):
        shape = res.shape
        K = a.shape[-1]
        a2 = a.reshape(-1, K)
        r2 = res.reshape(-1, shape[-1])
        if w16 is None:
            w16 = weight.to(torch.bfloat16)
        out = G.gemm(a2, w16, epilogue=G.EPI_RES, bias=bias, res=r2, p=p, seed=seed, seed_t=seed_t)
        ctx.save_for_backward(a2, w16, seed_t)
        ctx.p, ctx.seed, ctx.a_shape = p, seed, a.shape
        return out.view(shape)

    @staticmethod
    def backward(ctx, g):
        a2, w16, seed_t = ctx.saved_tensors
        N = w16.shape[0]
        dy, db = _take_dz(g) or _drop_mask_bf16(g.reshape(-1, N), ctx.p, ctx.seed, seed_t, want_db=True)
        da = G.gemm(dy, w16, b_kn=True).view(ctx.a_shape) if ctx.needs_input_grad[0] else None
        dw, db2 = G.gemm_wgrad(dy, a2, want_db=db is None)
        return da, g, None, dw, (db if db is not None else db2), None, None, None


def proj_res(a, res, lin: torch.nn.Linear, p: float, w16=None):
    seed, seed_t = _ep_seed(p)
    with torch.autocast("cuda", enabled=False):
        return _register_tail(_ProjRes.apply(a, res, w16, lin.weight, lin.bias, float(p), seed, seed_t), p, seed, seed_t)


class _LnFfn(torch.autograd.Function):
    """out = x + dropout(W2 dropout(gelu_tanh(W1 LN(x) + b1)) + b2)  (detr/model.py:224,182,405-411): two launches forward
    (LayerNorm-prologue GEMM with the GELU epilogue; GEMM with the residual epilogue)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, w1_16, w2_16, w1, b1, w2, b2, p1, p2, seed1, seed2, seed_t, tail):
        B, L, C = x.shape
        x2 = x.reshape(B * L, C)
        if x2.stride(1) != 1 or x2.stride(0) % 8 or x2.data_ptr() % 16:
            x2 = x2.contiguous()
        if w1_16 is None:
            w1_16 = w1.to(torch.bfloat16)
        if w2_16 is None:
            w2_16 = w2.to(torch.bfloat16)
        g32, be32 = gamma.float(), beta.float()
        F_ = w1_16.shape[0]
        y1 = torch.empty(B * L, F_, dtype=torch.bfloat16, device=x.device)
        h, a_plain, _, stats = G.gemm_ln(x2, g32, be32, eps, w1_16, epilogue=G.EPI_GELU, bias=b1, aux=y1, p=p1, seed=seed1, seed_t=seed_t)
        out = G.gemm(h, w2_16, epilogue=G.EPI_RES, bias=b2, res=x2, p=p2, seed=seed2, seed_t=seed_t)
        ctx.save_for_backward(x2, g32, stats, a_plain, y1, h, w1_16, w2_16, seed_t)
        ctx.cfg = (p1, p2, seed1, seed2, (B, L, C))
        ctx.tail = tail if x2.data_ptr() == x.data_ptr() else None
        return out.view(B, L, C)

    @staticmethod
    def backward(ctx, g):
        x2, g32, stats, a_plain, y1, h, w1_16, w2_16, seed_t = ctx.saved_tensors
        p1, p2, seed1, seed2, (B, L, C) = ctx.cfg
        g2 = g.reshape(B * L, C)
        dy2, db2 = _take_dz(g) or _drop_mask_bf16(g2, p2, seed2, seed_t, want_db=True)
        dw2, db2b = G.gemm_wgrad(dy2, h, want_db=db2 is None)
        # through the second projection, its dropout and the GELU: one GEMM with the GELU-backward epilogue
        dh = G.gemm(dy2, w2_16, b_kn=True, epilogue=G.EPI_GELU_BWD, aux=y1, p=p1, seed=seed1, seed_t=seed_t)
        dw1, db1 = G.gemm_wgrad(dh, a_plain)
        da = G.gemm(dh, w1_16, b_kn=True)
        dx, dgam, dbet = _ln_backward(da, None, g2, x2, g32, stats, ctx.tail)
        return (_as_grad(dx, (B, L, C)), dgam, dbet, None, None, None, dw1, db1, dw2, (db2 if db2 is not None else db2b),
                None, None, None, None, None, None)


def ln_ffn(x, norm: torch.nn.LayerNorm, fc1: torch.nn.Linear, fc2: torch.nn.Linear, p1: float, p2: float, w1_16=None, w2_16=None):
    seed1, seed_t = _ep_seed(p1)
    seed2, seed_t2 = _ep_seed(p2)
    st = seed_t if seed_t is not None else seed_t2
    with torch.autocast("cuda", enabled=False):
        out = _LnFfn.apply(x, norm.weight, norm.bias, float(norm.eps), w1_16, w2_16, fc1.weight, fc1.bias, fc2.weight, fc2.bias,
                           float(p1), float(p2), seed1, seed2, st, _claim_tail(x))
    return _register_tail(out, p2, seed2, st)


class _Proj(torch.autograd.Function):
    """n projections of the SAME input as one GEMM: outs[i] (.., N_i) bf16 = a @ W_i^T + b_i, returned as n column-slice views of
    one buffer (the attention kernels take strided views as they are).  Used for the decoder's cross-attention key / value
    projections of the encoder memory for ALL layers (detr/model.py:179-180 executed once instead of six times).  The slices are
    separate OUTPUTS of this node: backward concatenates the n incoming gradients once -- slicing one output tensor outside
    would make autograd zero-fill, copy and add a full-width buffer per slice (measured: 0.25 ms per step)."""

    @staticmethod
    def forward(ctx, a, w16, b32, n_w, *params):
        weights, biases = params[:n_w], params[n_w:]
        K = a.shape[-1]
        a2 = a.reshape(-1, K)
        if w16 is None:
            w16 = (weights[0] if n_w == 1 else torch.cat(weights, 0)).to(torch.bfloat16)
        if b32 is None:
            b32 = (biases[0] if n_w == 1 else torch.cat(biases, 0)).float()
        a16 = a2 if a2.dtype == torch.bfloat16 else a2.to(torch.bfloat16)
        out = G.gemm(a16, w16, bias=b32).view(*a.shape[:-1], -1)
        ctx.save_for_backward(a16, w16)
        ctx.a_shape, ctx.a_dtype, ctx.splits = a.shape, a.dtype, [w.shape[0] for w in weights]
        return tuple(torch.split(out, ctx.splits, dim=-1))

    @staticmethod
    def backward(ctx, *grads):
        a16, w16 = ctx.saved_tensors
        M = a16.shape[0]
        gs = [g.reshape(M, n) if g is not None else a16.new_zeros(M, n) for g, n in zip(grads, ctx.splits)]
        g2 = gs[0] if len(gs) == 1 else torch.cat(gs, dim=1)
        if g2.dtype != torch.bfloat16:
            g2 = g2.to(torch.bfloat16)
        # the input gradient leaves the GEMM in the input's dtype (fp32 for the encoder memory): no cast kernel behind it
        odt = ctx.a_dtype if ctx.a_dtype in (torch.float32, torch.bfloat16) else torch.bfloat16
        da = G.gemm(g2, w16, b_kn=True, out_dtype=odt).view(ctx.a_shape).to(ctx.a_dtype) if ctx.needs_input_grad[0] else None
        dw, db = G.gemm_wgrad(g2, a16)
        return (da, None, None, None, *_split_rows(dw, ctx.splits), *_split_rows(db, ctx.splits))


def proj(a, linears, w16=None, b32=None):
    """-> tuple of len(linears) projections of `a` (column slices of one GEMM's output)."""
    with torch.autocast("cuda", enabled=False):
        return _Proj.apply(a, w16, b32, len(linears), *[l.weight for l in linears], *[l.bias for l in linears])
