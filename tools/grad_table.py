#!/usr/bin/env python
"""Diagnostic: per-parameter gradient error of Encoder + Decoder at the BASELINE shape against the fp32 CPU oracle, next to the
error of the reference-style bf16-autocast oracle (the table behind tests/test_gpu_model_fullsize.py's gate).
    python tools/grad_table.py [--q 100] [--fp32]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "detr-object-detection_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from oracle import detr_oracle as O  # noqa: E402
import test_gpu_model_fullsize as T  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--q", type=int, default=100)
    ap.add_argument("--fp32", action="store_true")
    args = ap.parse_args()
    from detr_b200.harness import positional_encoding_tokens
    from detr_b200.model import DETRConfig, Decoder, Encoder
    cuda = torch.device("cuda:0")
    torch.manual_seed(11)
    Q = args.q
    cfg = DETRConfig(num_classes=91, num_object_queries=Q)
    enc, dec = Encoder(cfg).eval(), Decoder(cfg).eval()
    with torch.no_grad():
        for p in list(enc.parameters()) + list(dec.parameters()):
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    B, eh, ew = 2, 25, 34
    heights, widths = torch.tensor([800, 640], dtype=torch.int32), torch.tensor([1066, 900], dtype=torch.int32)
    x = torch.randn(B, eh * ew, 256)
    qe = 0.5 * torch.randn(Q, 256)
    pos = O.positional_encoding(eh, ew, heights, widths).flatten(2).permute(0, 2, 1).contiguous()
    mask = O.padding_mask(eh, ew, heights, widths).flatten(1)
    w = torch.randn(B, 6, Q, 256)
    mem_r, out_r, gx_r, gp_r = T._run_oracle(enc, dec, x, pos, qe, mask, w, autocast=False)
    mem_b, out_b, gx_b, gp_b = T._run_oracle(enc, dec, x, pos, qe, mask, w, autocast=True)
    enc, dec = enc.to(cuda), dec.to(cuda)
    for p in list(enc.parameters()) + list(dec.parameters()):
        p.grad = None
    pos_d, mask_d = positional_encoding_tokens(eh, ew, heights.to(cuda), widths.to(cuda), 32, 128, 10000)
    xg = x.to(cuda).requires_grad_(True)
    qe_p = qe.to(cuda).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not args.fp32):
        mem = enc(xg, position_embedding=pos_d, key_padding_mask=mask_d)
        out = dec(mem, position_embedding=pos_d, object_query_embedding=qe_p[None].expand(B, -1, -1), key_padding_mask=mask_d)
    (out.float() * w.to(cuda)).sum().backward()
    rows = [("memory", mem, mem_r, mem_b), ("decoded", out, out_r, out_b), ("grad_x", xg.grad, gx_r, gx_b),
            ("query_embed", qe_p.grad, gp_r["query_embed"], gp_b["query_embed"])]
    for prefix, mod in (("enc.", enc), ("dec.", dec)):
        for n, p in mod.named_parameters():
            rows.append((prefix + n, p.grad, gp_r[prefix + n], gp_b[prefix + n]))
    print(f"{'tensor':58s} {'scale':>10s} {'err':>10s} {'err_ref16':>10s} {'err/scale':>9s} {'ref/scale':>9s} {'err/ref':>8s}")
    for name, got, ref, ref16 in rows:
        got = got.detach().float().cpu()
        scale = ref.abs().max().item() + 1e-20
        err = (got - ref).abs().max().item()
        er = (ref16.float() - ref).abs().max().item()
        print(f"{name:58s} {scale:10.3e} {err:10.3e} {er:10.3e} {err / scale:9.2e} {er / scale:9.2e} {err / (er + 1e-20):8.2f}")


if __name__ == "__main__":
    main()
