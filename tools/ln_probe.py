#!/usr/bin/env python
"""Timing probe of the LayerNorm-prologue GEMM (csrc/gemm.cu gemm_ln_kernel): which part of the launch costs what.
Variants at the encoder q|k|v and FFN1 shapes, L2 warm (back-to-back launches) and cold (flushed)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "detr-object-detection_b200"))
import torch  # noqa: E402

from detr_b200 import gemm as G  # noqa: E402

dev = torch.device("cuda:0")
flushbuf = torch.ones(64 * 1024 * 1024, device=dev)


def timeit(fn, iters=20, flush=False):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flushbuf.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def burst(fn, n=50):
    """n launches back to back between two events: per-launch time with launch latency overlapped (as in a CUDA graph)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n


for M in (6800, 800):
    x16 = torch.randn(M, 256, device=dev).bfloat16(); x32 = x16.float()
    gam = torch.ones(256, device=dev); bet = torch.zeros(256, device=dev)
    pos = torch.randn(M, 256, device=dev)
    for N, npe in ((768, 512), (768, 0), (256, 0), (256, 256), (2048, 0)):
        w = (torch.randn(N, 256, device=dev) * 0.05).bfloat16(); bias = torch.randn(N, device=dev)
        aux = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for name, fn in (
            ("ln bf16x", lambda: G.gemm_ln(x16, gam, bet, 1e-5, w, addend=pos if npe else None, rows_per_batch=M, add_sb=0, add_sr=256, n_pos_end=npe, bias=bias)),
            ("ln bf16x noside", lambda: G.gemm_ln(x16, gam, bet, 1e-5, w, addend=pos if npe else None, rows_per_batch=M, add_sb=0, add_sr=256, n_pos_end=npe, bias=bias, want_operands=False)),
            ("ln fp32x", lambda: G.gemm_ln(x32, gam, bet, 1e-5, w, addend=pos if npe else None, rows_per_batch=M, add_sb=0, add_sr=256, n_pos_end=npe, bias=bias)),
            ("ln bf16x gelu", lambda: G.gemm_ln(x16, gam, bet, 1e-5, w, epilogue=G.EPI_GELU, bias=bias, aux=aux, p=0.1, seed=1) if npe == 0 else None),
            ("stream bias", lambda: G.gemm(x16, w, bias=bias)),
            ("stream gelu", lambda: G.gemm(x16, w, epilogue=G.EPI_GELU, bias=bias, aux=aux, p=0.1, seed=1)),
            ("cublaslt", lambda: torch.nn.functional.linear(x16, w, bias.bfloat16())),
        ):
            if name == "ln bf16x gelu" and npe:
                continue
            print(f"M={M:5d} N={N:4d} npe={npe:3d} {name:18s} single warm {timeit(fn):7.1f} us  cold {timeit(fn, flush=True):7.1f} us  burst {burst(fn):7.1f} us")
