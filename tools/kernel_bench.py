#!/usr/bin/env python
"""Stand-alone timing of every kernel of libdetr_b200.so at the BASELINE shapes (CUDA events on the launching
stream, >=3 warm-ups, L2 flushed between timed iterations: --flush read (default) streams a 256 MB buffer through L2
with loads, leaving it full of CLEAN foreign lines; --flush write fills the buffer, which leaves up to 126 MB of DIRTY
lines whose write-back is then charged to the kernel under test -- tens of microseconds for the HBM-bound kernels).  Prints one JSON line per
kernel with its algorithmic work and roofline fraction (peaks: MEASURED_PEAKS.json, burst figure -- kernels timed
alone).  Also the command `ncu` is pointed at (tools/kernel_bench.py --only attention --iters 1)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "detr-object-detection_b200"))
import torch  # noqa: E402

from detr_b200 import HungarianMatcher, SetCriterion, _lib, pack_targets  # noqa: E402
from detr_b200.attention import attention_backward, attention_forward  # noqa: E402


def synth_predictions(batch, layers, queries, num_classes, seed):
    """SURVEY.md 8(d) config 3 inputs (same generator as the test fixtures use)."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(batch, layers, queries, num_classes + 1, generator=g)
    boxes = torch.randn(batch, layers, queries, 4, generator=g).sigmoid()
    return logits, boxes


def synth_targets(batch, max_gt, num_classes, seed, min_gt=1):
    g = torch.Generator().manual_seed(seed)
    n = torch.randint(min_gt, max_gt + 1, (batch,), generator=g).tolist()
    labels, boxes = [], []
    for m in n:
        c = torch.rand(m, 2, generator=g) * 0.6 + 0.2
        s = torch.rand(m, 2, generator=g) * 0.30 + 0.02
        boxes.append(torch.cat([c - s / 2, c + s / 2], dim=1).float())
        labels.append(torch.randint(0, num_classes, (m,), generator=g, dtype=torch.int64))
    return labels, boxes


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p["bf16_tflops"], p["hbm_gbs"], "measured burst"
    except Exception:
        return 1590.0, 6650.0, "fallback"


class Flush:
    def __init__(self, dev, mode):
        self.buf = torch.ones(64 * 1024 * 1024, device=dev)
        self.mode = mode

    def __call__(self):
        if self.mode == "write":
            self.buf.fill_(1.0)
        else:
            self.sink = self.buf.sum()


WARMUP = 3


def timeit(fn, iters, flush):
    for _ in range(WARMUP):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3, help="untimed calls before each measurement (0 under ncu: one launch per shape)")
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--flush", choices=("read", "write"), default="read")
    args = ap.parse_args()
    global WARMUP
    WARMUP = args.warmup
    dev = torch.device("cuda:0")
    tf, hbm, src = peaks()
    flush = Flush(dev, args.flush)
    out = []

    def attn(name, B, nh, L, S):
        C = nh * 32
        q = torch.randn(B, L, C, device=dev).bfloat16(); k = torch.randn(B, S, C, device=dev).bfloat16()
        v = torch.randn(B, S, C, device=dev).bfloat16(); do = torch.randn(B, L, C, device=dev).bfloat16()
        o, lse = attention_forward(q, k, v, dropout_p=args.dropout, seed=1)
        f_ms = timeit(lambda: attention_forward(q, k, v, dropout_p=args.dropout, seed=1), args.iters, flush)
        b_ms = timeit(lambda: attention_backward(do, q, k, v, o, lse, dropout_p=args.dropout, seed=1), args.iters, flush)
        fl = 4.0 * L * S * C * B
        byt = (2 * C * (2 * L + 2 * S) + 4 * nh * L) * B
        for tag, ms, f in (("fwd", f_ms, fl), ("bwd", b_ms, 2.5 * fl)):
            out.append({"kernel": f"attention_{tag} {name}", "shape": [B, nh, L, S], "ms": round(ms, 4), "flops": f, "bytes_fwd_algorithmic": byt,
                        "achieved_tflops": round(f / ms / 1e9, 2), "frac_of_tensor_peak": round(f / ms / 1e9 / tf, 4), "bound": "tensor (MUFU-limited at d=32)",
                        "dropout_p": args.dropout})

    if not args.only or "attention" in args.only:
        attn("encoder self (config 2)", 8, 8, 850, 850)
        attn("decoder cross (config 2)", 8, 8, 100, 850)
        attn("decoder self (config 2)", 8, 8, 100, 100)
        attn("encoder self DC5 (config 4)", 2, 8, 3350, 3350)
        attn("decoder self Q=300 (config 5)", 4, 8, 300, 300)
        attn("decoder cross Q=300 (config 5)", 4, 8, 300, 850)

    if not args.only or "matcher" in args.only:
        for (B, L, maxm, tag) in ((256, 6, 100, "config 3"), (8, 6, 20, "config 2")):
            Q, NC = 100, 91
            logits, boxes = synth_predictions(B, L, Q, NC, seed=0)
            labels, gts = synth_targets(B, maxm, NC, seed=1)
            logits, boxes = logits.to(dev), boxes.to(dev)
            pt = pack_targets([l.to(dev) for l in labels], [g.to(dev) for g in gts], Q, dev)
            m = HungarianMatcher(1.0, 5.0, 2.0)
            ms = timeit(lambda: m.match_layers(logits, boxes, pt), args.iters, flush)
            byt = sum(38400 + 40 * c for c in pt.counts) * L
            out.append({"kernel": f"hungarian_match (cost + assignment fused) {tag}", "shape": [B, L, Q, NC + 1], "sum_gt": pt.total, "ms": round(ms, 4),
                        "bytes_algorithmic": byt, "achieved_gbs": round(byt / ms / 1e6, 2), "frac_of_hbm_peak": round(byt / ms / 1e6 / hbm, 5),
                        "images_per_s": round(B / ms * 1e3, 1), "bound": "latency (serial augmenting paths), reported against HBM for transparency"})
            crit = SetCriterion(NC, m).to(dev)
            lg = logits.clone().requires_grad_(True); bx = boxes.clone().requires_grad_(True)
            tg = {"class_idx": [l.to(dev) for l in labels], "boxes_normalized": [g.to(dev) for g in gts]}
            def crit_step():
                flush()
                loss = sum(v for k, v in crit({"pred_logits": lg, "pred_boxes": bx}, tg).items() if k.startswith("loss"))
                flush()
                loss.backward()
            for _ in range(WARMUP):
                crit_step()
            with _lib.profile() as prof:
                for _ in range(args.iters):
                    crit_step()
            torch.cuda.synchronize()
            for (n, _), ms_c in prof.median().items():
                if "criterion" in n:
                    byt_c = sum((36800 + 40 * cc) if "fwd" in n else (36800 * 2 + 1600 + 40 * cc) for cc in pt.counts) * L
                    out.append({"kernel": f"{n} {tag}", "ms": round(ms_c, 4), "bytes_algorithmic": byt_c, "achieved_gbs": round(byt_c / ms_c / 1e6, 2),
                                "frac_of_hbm_peak": round(byt_c / ms_c / 1e6 / hbm, 5), "bound": "hbm"})
    if not args.only or "gemm" in args.only:
        from detr_b200 import gemm as G
        import torch.nn.functional as F

        def gemm_row(name, ms, fl, byt, extra=None):
            r = {"kernel": name, "ms": round(ms, 4), "flops": fl, "bytes_algorithmic": byt, "achieved_tflops": round(fl / ms / 1e9, 2),
                 "frac_of_tensor_peak": round(fl / ms / 1e9 / tf, 4), "achieved_gbs": round(byt / ms / 1e6, 1),
                 "frac_of_hbm_peak": round(byt / ms / 1e6 / hbm, 4), "bound": "tensor / hbm (whichever fraction is larger)"}
            r.update(extra or {})
            out.append(r)

        for (M, N, K, tag) in ((6800, 768, 256, "enc q|k|v"), (6800, 256, 256, "enc out-proj"), (6800, 2048, 256, "enc FFN1"),
                               (6800, 256, 2048, "enc FFN2"), (800, 768, 256, "dec q|k|v"), (800, 2048, 256, "dec FFN1"),
                               (6800, 1536, 256, "dec cross K stacked"), (26800, 2048, 256, "DC5 FFN1")):
            a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
            bias = torch.randn(N, device=dev); b16 = bias.bfloat16()
            fl = 2.0 * M * N * K
            byt = 2.0 * (M * K + N * K + M * N)
            ms = timeit(lambda: G.gemm(a, w, bias=bias), args.iters, flush)
            gemm_row(f"gemm_stream bias {tag}", ms, fl, byt, {"shape": [M, N, K]})
            ms = timeit(lambda: F.linear(a, w, b16), args.iters, flush)
            gemm_row(f"COMPARATOR cuBLASLt F.linear {tag}", ms, fl, byt, {"shape": [M, N, K]})
            if K == 256:
                x = torch.randn(M, 256, device=dev).bfloat16(); gam = torch.ones(256, device=dev); bet = torch.zeros(256, device=dev)
                pos = torch.randn(M, 256, device=dev)
                npe = 512 if N == 768 else 0
                ms = timeit(lambda: G.gemm_ln(x, gam, bet, 1e-5, w, addend=pos if npe else None, rows_per_batch=M, add_sb=0, add_sr=256,
                                              n_pos_end=npe, bias=bias), args.iters, flush)
                gemm_row(f"gemm_ln bias {tag}", ms, fl, byt + (4.0 * M * 256 if npe else 0) + 2.0 * M * 256 * (2 if npe else 1), {"shape": [M, N, K]})
                if N == 2048:
                    aux = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
                    ms = timeit(lambda: G.gemm_ln(x, gam, bet, 1e-5, w, epilogue=G.EPI_GELU, bias=bias, aux=aux, p=args.dropout, seed=1),
                                args.iters, flush)
                    gemm_row(f"gemm_ln GELU+dropout {tag}", ms, fl, byt + 2.0 * M * N + 2.0 * M * 256, {"shape": [M, N, K]})
                    ms = timeit(lambda: F.gelu(F.linear(F.layer_norm(x.float(), (256,), gam, bet).bfloat16(), w, b16), approximate="tanh"),
                                args.iters, flush)
                    gemm_row(f"COMPARATOR ATen LN + cuBLASLt + ATen GELU {tag}", ms, fl, byt + 2.0 * M * N + 2.0 * M * 256, {"shape": [M, N, K]})
            if N == 256:
                res = torch.randn(M, N, device=dev).bfloat16()
                ms = timeit(lambda: G.gemm(a, w, epilogue=G.EPI_RES, bias=bias, res=res, p=args.dropout, seed=1), args.iters, flush)
                gemm_row(f"gemm_stream dropout+residual {tag}", ms, fl, byt + 2.0 * M * N, {"shape": [M, N, K]})
            dy = (torch.randn(M, N, device=dev) * 0.1).bfloat16()
            ms = timeit(lambda: G.gemm(dy, w, b_kn=True), args.iters, flush)
            gemm_row(f"gemm_stream dgrad {tag}", ms, fl, 2.0 * (M * N + N * K + M * K), {"shape": [M, K, N]})
            ms = timeit(lambda: dy @ w, args.iters, flush)
            gemm_row(f"COMPARATOR cuBLASLt dgrad {tag}", ms, fl, 2.0 * (M * N + N * K + M * K), {"shape": [M, K, N]})
            ms = timeit(lambda: G.gemm_wgrad(dy, a), args.iters, flush)
            gemm_row(f"gemm_wgrad (+bias grad) {tag}", ms, fl, 2.0 * (M * N + M * K) + 4.0 * N * K, {"shape": [N, K, M]})
            ms = timeit(lambda: (torch.mm(dy.t(), a, out_dtype=torch.float32), dy.float().sum(0)), args.iters, flush)
            gemm_row(f"COMPARATOR cuBLASLt wgrad + ATen colsum {tag}", ms, fl, 2.0 * (M * N + M * K) + 4.0 * N * K, {"shape": [N, K, M]})

    if "sdpa" in args.only:
        # comparators named by SURVEY.md 2a: the strongest existing attention kernels on this box, same shapes, same L2 policy
        import torch.nn.functional as F
        from torch.nn.attention import SDPBackend, sdpa_kernel
        try:
            from flash_attn import flash_attn_func
        except Exception:
            flash_attn_func = None
        for (name, B, nh, L, S) in (("encoder self (config 2)", 8, 8, 850, 850), ("decoder cross (config 2)", 8, 8, 100, 850),
                                    ("decoder self (config 2)", 8, 8, 100, 100), ("encoder self DC5 (config 4)", 2, 8, 3350, 3350),
                                    ("decoder self Q=300 (config 5)", 4, 8, 300, 300), ("decoder cross Q=300 (config 5)", 4, 8, 300, 850)):
            fl = 4.0 * L * S * nh * 32 * B
            q = torch.randn(B, nh, L, 32, device=dev).bfloat16().requires_grad_(True)
            k = torch.randn(B, nh, S, 32, device=dev).bfloat16().requires_grad_(True)
            v = torch.randn(B, nh, S, 32, device=dev).bfloat16().requires_grad_(True)
            do = torch.randn(B, nh, L, 32, device=dev).bfloat16()
            for bname, backend in (("flash", SDPBackend.FLASH_ATTENTION), ("efficient", SDPBackend.EFFICIENT_ATTENTION), ("cudnn", SDPBackend.CUDNN_ATTENTION),
                                   ("math", SDPBackend.MATH)):
                try:
                    with sdpa_kernel(backend):
                        f = lambda: F.scaled_dot_product_attention(q, k, v, dropout_p=args.dropout)
                        o = f()
                        f_ms = timeit(f, args.iters, flush)
                        b_ms = timeit(lambda: torch.autograd.grad(o, (q, k, v), do, retain_graph=True), args.iters, flush)
                except Exception as ex:  # backend not available for this shape / build
                    out.append({"kernel": f"COMPARATOR sdpa[{bname}] {name}", "unavailable": str(ex)[:120]})
                    continue
                for tag, ms, fw in (("fwd", f_ms, fl), ("bwd", b_ms, 2.5 * fl)):
                    out.append({"kernel": f"COMPARATOR sdpa[{bname}] {tag} {name}", "shape": [B, nh, L, S], "ms": round(ms, 4), "flops": fw,
                                "achieved_tflops": round(fw / ms / 1e9, 2), "frac_of_tensor_peak": round(fw / ms / 1e9 / tf, 4), "dropout_p": args.dropout})
            if flash_attn_func is not None:
                try:
                    q2, k2, v2 = (t.detach().transpose(1, 2).contiguous().requires_grad_(True) for t in (q, k, v))
                    do2 = do.transpose(1, 2).contiguous()
                    f = lambda: flash_attn_func(q2, k2, v2, dropout_p=args.dropout)
                    o = f()
                    f_ms = timeit(f, args.iters, flush)
                    b_ms = timeit(lambda: torch.autograd.grad(o, (q2, k2, v2), do2, retain_graph=True), args.iters, flush)
                    for tag, ms, fw in (("fwd", f_ms, fl), ("bwd", b_ms, 2.5 * fl)):
                        out.append({"kernel": f"COMPARATOR flash_attn 2.8 {tag} {name}", "shape": [B, nh, L, S], "ms": round(ms, 4), "flops": fw,
                                    "achieved_tflops": round(fw / ms / 1e9, 2), "frac_of_tensor_peak": round(fw / ms / 1e9 / tf, 4), "dropout_p": args.dropout})
                except Exception as ex:
                    out.append({"kernel": f"COMPARATOR flash_attn {name}", "unavailable": str(ex)[:120]})

    for r in out:
        r["peak_source"] = src
        r["l2_flush"] = args.flush
        print(json.dumps(r))


if __name__ == "__main__":
    main()
