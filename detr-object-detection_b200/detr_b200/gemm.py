"""Raw launches of the tcgen05 GEMM kernels (csrc/gemm.cu): the nn.Linear layers of the transformer blocks
(detr/model.py:312-314,354,405-411) with LayerNorm / bias / GELU / dropout / residual fused into prologue and epilogue.

Everything here is 2-D: activations are (M, features) views of the (B, L, C) tensors."""
from __future__ import annotations

import contextlib
import os
from typing import Optional

import torch

from . import _lib

EPI_BIAS, EPI_GELU, EPI_RES, EPI_GELU_BWD, EPI_SIGMOID = 0, 1, 2, 3, 4
_DT = {torch.float32: 0, torch.bfloat16: 1}


def _mat(t: torch.Tensor, what: str) -> torch.Tensor:
    """bf16 2-D operand with unit column stride, 16-byte aligned rows (what TMA needs); copies only when it must."""
    if t.dim() != 2:
        raise ValueError(f"{what}: expected a 2-D matrix, got {tuple(t.shape)}")
    if t.dtype != torch.bfloat16:
        t = t.to(torch.bfloat16)
    if t.stride(1) != 1 or t.stride(0) % 8 or t.data_ptr() % 16:
        t = t.contiguous()
    return t


def gemm(a: torch.Tensor, b: torch.Tensor, *, b_kn: bool = False, epilogue: int = EPI_BIAS, bias: Optional[torch.Tensor] = None,
         out_dtype=torch.bfloat16, out: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
         res: Optional[torch.Tensor] = None, p: float = 0.0, seed: int = 0, seed_t: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out (M, N) = epilogue(a (M, K) @ b^T) with b (N, K), or a @ b with b (K, N) when `b_kn`.  See detr_gemm_bf16."""
    _lib.require_cuda(a, "gemm")
    a, b = _mat(a, "gemm(a)"), _mat(b, "gemm(b)")
    M, K = a.shape
    N = b.shape[1] if b_kn else b.shape[0]
    if (b.shape[0] if b_kn else b.shape[1]) != K:
        raise ValueError(f"gemm: inner sizes differ: a {tuple(a.shape)}, b {tuple(b.shape)}, b_kn={b_kn}")
    if epilogue == EPI_RES:
        out_dtype = res.dtype
        if res.stride(1) != 1 or res.stride(0) % 8 or res.data_ptr() % 16:
            res = res.contiguous()
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a.device)
    if bias is not None and bias.dtype != torch.float32:
        bias = bias.float()
    _lib.call("detr_gemm_bf16", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), int(b_kn), M, N, K, epilogue, _lib.ptr(bias),
              out.data_ptr(), _DT[out.dtype], out.stride(0), _lib.ptr(aux), aux.stride(0) if aux is not None else 0,
              _lib.ptr(res), res.stride(0) if res is not None else 0, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, _lib.ptr(seed_t),
              _lib.stream_ptr(), tag=("gemm", epilogue, M, N, K))
    return out


def gemm_ln(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, w: torch.Tensor, *, addend: Optional[torch.Tensor] = None,
            rows_per_batch: int = 0, add_sb: int = 0, add_sr: int = 0, n_pos_end: int = 0, epilogue: int = EPI_BIAS,
            bias: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None, want_operands: bool = True,
            p: float = 0.0, seed: int = 0, seed_t: Optional[torch.Tensor] = None):
    """out (M, N) bf16 = epilogue((LN(x) [+ addend]) @ w^T), x (M, 256) fp32 / bf16, w (N, 256) bf16.  Output columns below
    `n_pos_end` use LN(x) + addend.  Returns (out, a_plain, a_pos, stats) -- the bf16 operands and (2, M) mean / rstd the backward
    pass needs (None when not produced).  See detr_gemm_ln_bf16."""
    _lib.require_cuda(x, "gemm_ln")
    M, C = x.shape
    if C != 256 or x.dtype not in _DT:
        raise ValueError("gemm_ln: x must be (M, 256) float32 / bfloat16")
    if x.stride(1) != 1 or x.stride(0) % 8 or x.data_ptr() % 16:
        x = x.contiguous()
    w = _mat(w, "gemm_ln(w)")
    N = w.shape[0]
    dev = x.device
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    has_pos = n_pos_end > 0
    has_plain = n_pos_end < N
    a_plain = torch.empty(M, C, dtype=torch.bfloat16, device=dev) if (want_operands and has_plain) else None
    a_pos = torch.empty(M, C, dtype=torch.bfloat16, device=dev) if (want_operands and has_pos) else None
    stats = torch.empty(2, M, dtype=torch.float32, device=dev) if want_operands else None
    g32 = gamma if gamma.dtype == torch.float32 else gamma.float()
    b32 = beta if beta.dtype == torch.float32 else beta.float()
    if bias is not None and bias.dtype != torch.float32:
        bias = bias.float()
    _lib.call("detr_gemm_ln_bf16", x.data_ptr(), _DT[x.dtype], x.stride(0), g32.data_ptr(), b32.data_ptr(), float(eps),
              _lib.ptr(addend) if has_pos else None, add_sb, add_sr, rows_per_batch, n_pos_end, w.data_ptr(), w.stride(0), M, N, epilogue,
              _lib.ptr(bias), out.data_ptr(), out.stride(0), _lib.ptr(aux), aux.stride(0) if aux is not None else 0,
              _lib.ptr(a_plain), _lib.ptr(a_pos), stats[0].data_ptr() if stats is not None else None,
              stats[1].data_ptr() if stats is not None else None, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, _lib.ptr(seed_t),
              _lib.stream_ptr(), tag=("gemm_ln", epilogue, M, N))
    return out, a_plain, a_pos, stats


class _SideStream:
    """Weight gradients leave the critical path: inside a backward pass `gemm_wgrad` may launch on a second stream, forked from
    the caller's stream at the call and joined back ONCE, by an autograd engine callback at the end of the pass (under stream
    capture: a parallel branch of the graph).  The small decoder GEMMs occupy 4-32 of 148 SMs, so the branch runs beside the
    input-gradient chain instead of between its links.  Operands and scratch are allocated on the caller's stream and kept
    alive until the join.

    Contract (why this is opt-in, `wgrad_side_stream(True)`): nothing may READ a weight gradient before the backward pass ends.
    That excludes DistributedDataParallel's and the harness' bucketed all-reduce hooks, and it excludes every case in which
    autograd's AccumulateGrad launches a kernel of its own: `p.grad` must be None when backward starts (otherwise it adds in place),
    and the gradient tensor must be stealable -- sole owner, parameter's layout -- (otherwise it clones).  For that reason the
    OUTPUTS are not in the keep-alive list: a second reference turns the steal into a clone on the caller's stream, which under
    graph replay reads the buffer before the side stream has written it (found the hard way: tests/test_gpu_graph.py).
    `harness.GraphedTrainStep` meets the contract: it resets every `.grad` to None before each forward/backward and reads the
    gradients only after backward() has returned."""

    def __init__(self):
        self.enabled = os.environ.get("DETR_B200_WGRAD_STREAM", "0") == "1"
        self.streams = {}
        self.pending = {}          # graph task id -> [device, tensors kept alive until the join]

    def stream(self, dev: torch.device) -> "torch.cuda.Stream":
        st = self.streams.get(dev.index)
        if st is None:
            st = self.streams[dev.index] = torch.cuda.Stream(dev)
        return st

    def fork(self, dev: torch.device, keep):
        """Returns the side stream to launch on (already waiting for the caller's stream), or None outside a backward pass."""
        task = torch._C._current_graph_task_id()
        if not self.enabled or task < 0:
            return None
        side = self.stream(dev)
        entry = self.pending.get(task)
        if entry is None:
            # entries of passes that never reached their callback (backward raised): join them now so that the side stream never
            # stays forked and their tensors are released
            for stale in [t for t in self.pending if t != task]:
                self.join(stale)
            entry = self.pending[task] = [dev, []]
            torch.autograd.Variable._execution_engine.queue_callback(lambda: self.join(task))
        elif entry[0] != dev:
            raise RuntimeError("weight-gradient side stream: one process drives one GPU (backward pass spans several devices)")
        entry[1].append(keep)
        self.refork(dev)
        return side

    def refork(self, dev: torch.device) -> None:
        """Make the side stream wait for everything launched on the caller's stream so far."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        self.stream(dev).wait_event(ev)

    def join_now(self) -> None:
        """Inside a backward pass: make the caller's stream wait for everything launched on the side stream so far (for a node that
        consumes side-stream results before the pass ends, e.g. the unfold of the backbone's weight gradients)."""
        entry = self.pending.get(torch._C._current_graph_task_id())
        if entry is not None:
            ev = torch.cuda.Event()
            ev.record(self.stream(entry[0]))
            torch.cuda.current_stream(entry[0]).wait_event(ev)

    def debug_after_launch(self, dev) -> None:
        if os.environ.get("DETR_B200_WGRAD_STREAM_DEBUG", "") == "sync" and not torch.cuda.is_current_stream_capturing():
            self.stream(dev).synchronize()

    def join(self, task: int) -> None:
        if task not in self.pending:
            return
        dev, keep = self.pending.pop(task)
        ev = torch.cuda.Event()
        ev.record(self.stream(dev))
        torch.cuda.current_stream(dev).wait_event(ev)
        keep.clear()


_SIDE = _SideStream()


def wgrad_side_stream(enabled: bool) -> bool:
    """Switch the side-stream launch of weight-gradient GEMMs (see _SideStream); returns the previous setting."""
    prev, _SIDE.enabled = _SIDE.enabled, bool(enabled)
    return prev


def gemm_wgrad(dy: torch.Tensor, x0: torch.Tensor, x1: Optional[torch.Tensor] = None, n_switch: int = 0, want_db: bool = True):
    """(dw (N, K) fp32, db (N,) fp32 | None) = (dy^T @ x, column sums of dy); rows >= n_switch of dw use x1.  See detr_gemm_wgrad_bf16."""
    _lib.require_cuda(dy, "gemm_wgrad")
    dy, x0 = _mat(dy, "gemm_wgrad(dy)"), _mat(x0, "gemm_wgrad(x)")
    if x1 is not None:
        x1 = _mat(x1, "gemm_wgrad(x1)")
    M, N = dy.shape
    K = x0.shape[1]
    dev = dy.device
    dw = torch.empty(N, K, dtype=torch.float32, device=dev)
    db = torch.empty(N, dtype=torch.float32, device=dev) if want_db else None
    nws = _lib.load().detr_gemm_wgrad_workspace_floats(M, N, K)
    ws = torch.empty(nws, dtype=torch.float32, device=dev) if nws else None
    side = _SIDE.fork(dev, (dy, x0, x1, ws))       # operands and scratch only -- NOT the outputs, see _SideStream
    with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
        _lib.call("detr_gemm_wgrad_bf16", dy.data_ptr(), dy.stride(0), x0.data_ptr(), x0.stride(0), _lib.ptr(x1), x1.stride(0) if x1 is not None else 0,
                  n_switch if x1 is not None else N, M, N, K, dw.data_ptr(), _lib.ptr(db), _lib.ptr(ws), _lib.stream_ptr(),
                  tag=("wgrad", M, N, K), launches=2 if nws else 1)
    if side is not None:
        _SIDE.debug_after_launch(dev)
    return dw, db
