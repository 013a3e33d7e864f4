"""Row-wise operations of the transformer blocks on libdetr_b200 kernels."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib

_DT = {torch.float32: 0, torch.bfloat16: 1}


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
    """y = x W^T + b with cuBLASLt (bias in the GEMM epilogue) and a backward that (a) produces dW directly in fp32
    from the bf16 GEMM (no cast launch), (b) computes db with ONE pass of the colsum kernel instead of ATen's generic
    reduction.  `w16` / `b16` are bf16 shadows of the fp32 parameters (refreshed once per forward for the whole
    encoder / decoder by `ShadowedLinears.refresh`); when absent they are cast here.  `weights` / `biases` may hold
    several parameters that are stacked along the output dimension (the fused q|k projection)."""

    @staticmethod
    def forward(ctx, x, w16, b16, n_w, *params):
        weights, biases = params[:n_w], params[n_w:]
        if w16 is None:
            w16 = (weights[0] if n_w == 1 else torch.cat(weights, 0)).to(x.dtype)
        if b16 is None and biases:
            b16 = (biases[0] if len(biases) == 1 else torch.cat(biases, 0)).to(x.dtype)
        ctx.save_for_backward(x, w16)
        ctx.n_w, ctx.n_b = n_w, len(biases)
        ctx.splits = [w.shape[0] for w in weights]
        ctx.w_dtype = weights[0].dtype
        return F.linear(x, w16, b16)

    @staticmethod
    def backward(ctx, g):
        x, w16 = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        x2 = x.reshape(-1, x.shape[-1])
        dx = (g2 @ w16).view(x.shape) if ctx.needs_input_grad[0] else None
        dw = torch.mm(g2.t(), x2, out_dtype=torch.float32) if ctx.w_dtype == torch.float32 else (g2.t() @ x2).to(ctx.w_dtype)
        dws = (dw,) if ctx.n_w == 1 else torch.split(dw, ctx.splits, 0)
        dbs = ()
        if ctx.n_b:
            db = colsum(g2 if g2.stride(1) == 1 else g2.contiguous())
            dbs = (db,) if ctx.n_b == 1 else torch.split(db, ctx.splits, 0)
        return (dx, None, None, None, *dws, *dbs)


class _StackedLinearFn(torch.autograd.Function):
    """Several nn.Linear layers applied to the SAME input as one GEMM: y_l = x W_l^T + b_l for l = 0..n-1, returned as n
    tensors that are column slices of one (.., n*N) buffer (the attention kernels take the strided views as they are).
    Backward concatenates the n incoming gradients once and runs ONE dgrad GEMM, ONE wgrad GEMM and ONE colsum instead of n of
    each plus n-1 gradient accumulations on x.  Used for the decoder's cross-attention key / value projections, which read
    the same encoder memory in every layer (detr/model.py:179-180 executes them six times)."""

    @staticmethod
    def forward(ctx, x, w16, b16, n, *params):
        weights, biases = params[:n], params[n:]
        if w16 is None:
            w16 = torch.cat(weights, 0).to(x.dtype)
        if b16 is None:
            b16 = torch.cat(biases, 0).to(x.dtype)
        ctx.save_for_backward(x, w16)
        ctx.n, ctx.N, ctx.w_dtype = n, weights[0].shape[0], weights[0].dtype
        y = F.linear(x, w16, b16)
        return tuple(y[..., i * ctx.N:(i + 1) * ctx.N] for i in range(n))

    @staticmethod
    def backward(ctx, *grads):
        x, w16 = ctx.saved_tensors
        n, N = ctx.n, ctx.N
        gs = [g if g is not None else x.new_zeros(x.shape[:-1] + (N,)) for g in grads]
        g2 = torch.cat([g.reshape(-1, N) for g in gs], dim=1)          # (M, n*N) bf16
        x2 = x.reshape(-1, x.shape[-1])
        dx = (g2 @ w16).view(x.shape) if ctx.needs_input_grad[0] else None
        dw = torch.mm(g2.t(), x2, out_dtype=torch.float32) if ctx.w_dtype == torch.float32 else (g2.t() @ x2).to(ctx.w_dtype)
        db = colsum(g2)
        if ctx.w_dtype != torch.float32:
            db = db.to(ctx.w_dtype)
        return (dx, None, None, None, *torch.split(dw, N, 0), *torch.split(db, N, 0))


def stacked_linear(x: torch.Tensor, linears, w16=None, b16=None):
    """[lin(x) for lin in linears] as one GEMM under bf16 autocast (strided views of one buffer); plain loop otherwise."""
    if torch.is_autocast_enabled() and x.is_cuda and torch.get_autocast_dtype("cuda") == torch.bfloat16:
        with torch.autocast("cuda", enabled=False):
            return _StackedLinearFn.apply(x.to(torch.bfloat16), w16, b16, len(linears), *[l.weight for l in linears], *[l.bias for l in linears])
    return tuple(F.linear(x, l.weight, l.bias) for l in linears)


def linear(x: torch.Tensor, weight, bias, w16=None, b16=None) -> torch.Tensor:
    """nn.functional.linear under bf16 autocast through `_LinearFn`; plain F.linear otherwise (fp32 mode).
    `weight` / `bias` may be tuples of parameters stacked along the output dimension."""
    ws = weight if isinstance(weight, (tuple, list)) else (weight,)
    bs = () if bias is None else (bias if isinstance(bias, (tuple, list)) else (bias,))
    if torch.is_autocast_enabled() and x.is_cuda and torch.get_autocast_dtype("cuda") == torch.bfloat16:
        with torch.autocast("cuda", enabled=False):
            return _LinearFn.apply(x.to(torch.bfloat16), w16, b16, len(ws), *ws, *bs)
    w = ws[0] if len(ws) == 1 else torch.cat(tuple(ws), 0)
    b = None if not bs else (bs[0] if len(bs) == 1 else torch.cat(tuple(bs), 0))
    return F.linear(x, w, b)


def _ep_seed(p: float):
    """(host seed, device step tensor) for an epilogue dropout mask: host draw (follows torch.manual_seed, no sync) plus
    the device-side step counter registered for CUDA-graph replays (attention.set_dropout_step_tensor)."""
    if p <= 0.0:
        return 0, None
    from . import attention
    return int(torch.randint(0, 2 ** 62, (1,)).item()), attention._STEP_TENSOR


class _LinearEpilogue(torch.autograd.Function):
    """One pre-LN block tail on cuBLASLt + ONE elementwise kernel each way (csrc/rowops.cu, `detr_epilogue_*`):

        mode 0:  out = residual + dropout(x W^T + b)        attention output projection, second FFN projection
        mode 1:  out = dropout(gelu_tanh(x W^T + b))        first FFN projection

    Backward regenerates the dropout mask, produces dy in bf16 for the two GEMMs and the bias gradient in the same
    pass (no stored mask, no ATen dropout / masked_scale / gelu / gelu_backward / add launches, no separate colsum)."""

    @staticmethod
    def forward(ctx, mode, x, residual, w16, b16, weight, bias, p, seed, seed_t):
        if w16 is None:
            w16 = weight.to(torch.bfloat16)
        if b16 is None:
            b16 = bias.to(torch.bfloat16)
        y = F.linear(x, w16, b16)
        N = y.shape[-1]
        M = y.numel() // N
        if mode == 0:
            res = residual if residual.is_contiguous() else residual.contiguous()
            out = torch.empty_like(res)
        else:
            res = None
            out = torch.empty_like(y)
        _lib.call("detr_epilogue_fwd", mode, _lib.ptr(res), _DT[res.dtype] if res is not None else 1, y.data_ptr(), out.data_ptr(),
                  M, N, float(p), seed & 0xFFFFFFFFFFFFFFFF, _lib.ptr(seed_t), _lib.stream_ptr())
        ctx.save_for_backward(x, w16, y if mode == 1 else None, seed_t)
        ctx.mode, ctx.p, ctx.seed, ctx.MN = mode, p, seed, (M, N)
        ctx.w_dtype = weight.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        x, w16, y, seed_t = ctx.saved_tensors
        M, N = ctx.MN
        if g.dtype not in _DT:
            g = g.float()
        g = g if g.is_contiguous() else g.contiguous()
        dy = torch.empty(M, N, dtype=torch.bfloat16, device=g.device)
        chunks = _lib.load().detr_epilogue_chunks(M, N)
        partial = torch.empty(chunks * N, dtype=torch.float32, device=g.device)
        db = torch.empty(N, dtype=torch.float32, device=g.device)
        _lib.call("detr_epilogue_bwd", ctx.mode, g.data_ptr(), _DT[g.dtype], _lib.ptr(y), dy.data_ptr(), partial.data_ptr(), db.data_ptr(),
                  _lib.zero_counters(g.device).data_ptr(), M, N, float(ctx.p), ctx.seed & 0xFFFFFFFFFFFFFFFF, _lib.ptr(seed_t), _lib.stream_ptr())
        x2 = x.reshape(M, -1)
        dx = (dy @ w16).view(x.shape) if ctx.needs_input_grad[1] else None
        dw = torch.mm(dy.t(), x2, out_dtype=torch.float32) if ctx.w_dtype == torch.float32 else (dy.t() @ x2).to(ctx.w_dtype)
        d_res = g if (ctx.mode == 0 and ctx.needs_input_grad[2]) else None
        return None, dx, d_res, None, None, dw, db.to(ctx.w_dtype) if ctx.w_dtype != torch.float32 else db, None, None, None


def fused_epilogues_enabled(x: torch.Tensor) -> bool:
    return x.is_cuda and torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16


def linear_dropout_add(x, residual, lin: torch.nn.Linear, p: float, w16=None, b16=None) -> torch.Tensor:
    """residual + dropout(lin(x)) under bf16 autocast (result in the residual's dtype)."""
    seed, seed_t = _ep_seed(p)
    with torch.autocast("cuda", enabled=False):
        return _LinearEpilogue.apply(0, x.to(torch.bfloat16), residual, w16, b16, lin.weight, lin.bias, p, seed, seed_t)


def linear_gelu_dropout(x, lin: torch.nn.Linear, p: float, w16=None, b16=None) -> torch.Tensor:
    """dropout(gelu_tanh(lin(x))) under bf16 autocast (bf16 result)."""
    seed, seed_t = _ep_seed(p)
    with torch.autocast("cuda", enabled=False):
        return _LinearEpilogue.apply(1, x.to(torch.bfloat16), None, w16, b16, lin.weight, lin.bias, p, seed, seed_t)


class ShadowedLinears:
    """bf16 shadows of a module tree's Linear weights and biases, refreshed with ONE multi-tensor copy per forward
    (instead of one cast kernel per weight per use under autocast).  Groups of parameters that are used stacked
    (q|k of self-attention) share one shadow buffer, so no torch.cat is launched either."""

    def __init__(self):
        self.groups = []      # (key, [weights], [biases])
        self.shadow = {}      # key -> (w16, b16)
        self.shadow32 = {}    # key -> stacked fp32 bias
        self._src, self._dst = [], []
        self._device = None
        self._fresh = None    # event of a prefetch that the next refresh() only has to wait for

    def register(self, key, weights, biases, pad_rows=None):
        """`pad_rows[i]` >= weights[i].shape[0]: rows the i-th weight occupies in the stacked shadow (the rest stays zero) -- the
        prediction heads' 92- and 4-row weights padded to whole MMA tiles."""
        self.groups.append((key, list(weights), list(biases), list(pad_rows) if pad_rows else None))

    def _build(self, device):
        self.shadow.clear(); self._src, self._dst = [], []
        for key, ws, bs, pads in self.groups:
            rows = pads if pads else [w.shape[0] for w in ws]
            out = sum(rows)
            make = torch.zeros if pads else torch.empty
            w16 = make(out, ws[0].shape[1], dtype=torch.bfloat16, device=device)
            b16 = make(out, dtype=torch.bfloat16, device=device) if bs else None
            b32 = make(out, dtype=torch.float32, device=device) if bs else None   # stacked fp32 bias for the GEMM epilogues
            o = 0
            for i, w in enumerate(ws):
                self._src.append(w); self._dst.append(w16[o:o + w.shape[0]])
                if bs:
                    self._src.append(bs[i]); self._dst.append(b16[o:o + w.shape[0]])
                    self._src.append(bs[i]); self._dst.append(b32[o:o + w.shape[0]])
                o += rows[i]
            self.shadow[key] = (w16, b16)
            self.shadow32[key] = b32
        self._device = device

    @torch.no_grad()
    def refresh(self, device):
        """Bring the shadows up to date on the current stream -- or, if `prefetch` already did it on another stream, just make
        the current stream wait for that."""
        ev = self._fresh
        if ev is not None:
            self._fresh = None
            if self._device == device:
                torch.cuda.current_stream(device).wait_event(ev)
                return
        self._copy(device)

    @torch.no_grad()
    def prefetch(self, device):
        """Refresh on the CURRENT stream (the caller switched to a side stream) and leave an event for the `refresh` call of the
        consumer: the copies depend on the parameters only, so they can run beside whatever precedes the consumer (the ResNet
        forward in the full model)."""
        self._copy(device)
        self._fresh = torch.cuda.Event()
        self._fresh.record(torch.cuda.current_stream(device))

    def _copy(self, device):
        if self._device != device:
            self._build(device)
        # one multi-tensor launch per destination dtype (a list that mixes dtypes makes _foreach_copy_ fall back to one copy
        # kernel per tensor: ~400 launches per step)
        for dt in (torch.bfloat16, torch.float32):
            dst = [d for d in self._dst if d.dtype == dt]
            if dst:
                torch._foreach_copy_(dst, [p.detach() for p, d in zip(self._src, self._dst) if d.dtype == dt])

    def get(self, key):
        return self.shadow.get(key, (None, None))

    def get_w_b32(self, key):
        """(bf16 stacked weight, fp32 stacked bias) of a group, (None, None) before the first refresh."""
        return self.shadow.get(key, (None, None))[0], self.shadow32.get(key)
# ------------------------------------------------------------------------------------------------ fused LayerNorm


class _LayerNormAdd(torch.autograd.Function):
    """(y, y2) = (LN(x), LN(x) + addend) in one kernel; y / y2 in `out_dtype`.  `addend` is fp32 (B, R, C) (any batch /
    row strides, e.g. an expanded (R, C) embedding) or None; `want_y=False` skips writing y (decoder cross-attention
    only needs the query = LN(x) + query_embed).  `pass_x=True` adds a third output: x itself, to be used as the residual
    input of the block's tail (x + f(LN(x))); the gradient that comes back through it is added to dx inside the LayerNorm
    backward kernel instead of by a separate autograd accumulation kernel."""

    @staticmethod
    def forward(ctx, x, gamma, beta, addend, eps, out_dtype, want_y, pass_x=False):
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
        return (y, y2, x) if pass_x else (y, y2)

    @staticmethod
    def backward(ctx, dy, dy2, dres=None):
        x2, g32, stats = ctx.saved_tensors
        rows, C = x2.shape
        gs = [t for t in (dy, dy2) if t is not None]
        if not gs:
            return dres, None, None, None, None, None, None, None
        if dres is not None:
            dres = dres.to(x2.dtype).reshape(rows, C).contiguous()
        gdt = torch.float32 if any(t.dtype == torch.float32 for t in gs) else torch.bfloat16
        prep = lambda t: None if t is None else t.to(gdt).reshape(rows, C).contiguous()
        dyc, dy2c = prep(dy), prep(dy2)
        dx = torch.empty(rows, C, dtype=x2.dtype, device=x2.device)
        grid = _lib.load().detr_layernorm_grid(rows)
        partial = torch.empty(grid * 2 * C, dtype=torch.float32, device=x2.device)
        dgb = torch.empty(2, C, dtype=torch.float32, device=x2.device)
        _lib.call("detr_layernorm_bwd", _lib.ptr(dyc), _lib.ptr(dy2c), _DT[gdt], _lib.ptr(dres), x2.data_ptr(), _DT[x2.dtype], x2.stride(0),
                  g32.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), dx.data_ptr(), partial.data_ptr(),
                  dgb[0].data_ptr(), dgb[1].data_ptr(), _lib.zero_counters(x2.device).data_ptr(), rows, C, _lib.stream_ptr())
        d_add = dy2.to(ctx.addend_dtype) if (ctx.has_addend and dy2 is not None and ctx.needs_input_grad[3]) else None
        return dx.view(ctx.shape), dgb[0], dgb[1], d_add, None, None, None, None


def layer_norm_add(x, norm: torch.nn.LayerNorm, addend=None, want_y: bool = True, pass_x: bool = False):
    """-> (LN(x), LN(x) + addend[, x]).  bf16 outputs under bf16 autocast (what the following GEMMs consume), else x.dtype.
    With `pass_x` the third output is x, to be used as the block's residual input (see _LayerNormAdd)."""
    out_dtype = torch.bfloat16 if (torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16) else x.dtype
    if x.dtype not in _DT or out_dtype not in _DT:
        raise TypeError(f"layer_norm_add supports float32 / bfloat16, got {x.dtype}")
    _lib.require_cuda(x, "layer_norm_add")
    with torch.autocast("cuda", enabled=False):
        return _LayerNormAdd.apply(x, norm.weight, norm.bias, addend, norm.eps, out_dtype, want_y, pass_x)
