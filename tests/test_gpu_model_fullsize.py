"""End-to-end transformer parity at the BASELINE shape: default DETRConfig (C=256, 8 heads, 6+6 layers) on the 25 x 34 = 850
token map of an 800 x 1066 image (7 key tiles per attention item, split items, deferred finishes) with the reference's corner
padding mask ACTIVE, and the config-5 decoder (300 object queries).  Compared with the fp32 CPU oracle
(oracle/detr_oracle.py, pinned to the real reference by tests/golden/make_golden.py):

  * decoder output (B, 6, Q, 256), encoder memory;
  * the input gradient and EVERY parameter gradient (LayerNorm gamma / beta from the LayerNorm backward kernel, biases from the
    weight-gradient GEMM's column sums, the stacked q|k|v and cross-attention K / V weight gradients, ...).

Gate (VERDICT r1 next #1a): err <= 2 * err_ref_bf16 + eps * scale, where err_ref_bf16 is the error of the
reference-style path itself under bf16 autocast (the oracle under CPU autocast) against the same fp32 oracle -- i.e. this
implementation may be at most twice as far from fp32 as the reference's own bf16 training path is; eps = 2e-3 of the tensor's
max magnitude absorbs tensors whose reference error happens to be ~0 (e.g. biases behind a softmax-invariant path)."""
import pytest
import torch

from oracle import detr_oracle as O

pytestmark = pytest.mark.gpu


def _run_oracle(enc, dec, x, pos, qe, mask, w, autocast):
    esd, dsd = dict(enc.named_parameters()), dict(dec.named_parameters())
    for p in list(esd.values()) + list(dsd.values()):
        p.grad = None
    xr = x.clone().requires_grad_(True)
    qr = qe.clone().requires_grad_(True)
    B = x.shape[0]
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        mem = O.encoder(esd, xr, pos, mask, len(enc.layers), 8)
        out = O.decoder(dsd, mem, pos, qr[None].expand(B, -1, -1), mask, len(dec.layers), 8)
    (out.float() * w).sum().backward()
    grads = {"enc." + n: p.grad.clone() for n, p in esd.items()}
    grads.update({"dec." + n: p.grad.clone() for n, p in dsd.items()})
    grads["query_embed"] = qr.grad.clone()
    return mem.detach().float(), out.detach().float(), xr.grad.clone(), grads


@pytest.mark.parametrize("autocast", [True, False])
@pytest.mark.parametrize("Q", [100, 300])
def test_baseline_shape_all_gradients_vs_oracle(cuda, autocast, Q):
    from detr_b200.harness import positional_encoding_tokens
    from detr_b200.model import DETRConfig, Decoder, Encoder
    if Q == 300 and not autocast:
        pytest.skip("config 5 is a bf16 training configuration; the fp32 mode is covered at Q=100")
    torch.manual_seed(11)
    cfg = DETRConfig(num_classes=91, num_object_queries=Q)
    enc, dec = Encoder(cfg).eval(), Decoder(cfg).eval()
    with torch.no_grad():   # non-trivial LayerNorm affine parameters and biases (init is 1 / 0)
        for p in list(enc.parameters()) + list(dec.parameters()):
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    B, eh, ew = 2, 25, 34
    heights, widths = torch.tensor([800, 640], dtype=torch.int32), torch.tensor([1066, 900], dtype=torch.int32)
    x = torch.randn(B, eh * ew, 256)
    qe = 0.5 * torch.randn(Q, 256)
    pos = O.positional_encoding(eh, ew, heights, widths).flatten(2).permute(0, 2, 1).contiguous()
    mask = O.padding_mask(eh, ew, heights, widths).flatten(1)
    assert mask[1].any() and not mask[0].any()          # the second image has a masked corner
    w = torch.randn(B, 6, Q, 256)

    mem_r, out_r, gx_r, gp_r = _run_oracle(enc, dec, x, pos, qe, mask, w, autocast=False)
    mem_b, out_b, gx_b, gp_b = _run_oracle(enc, dec, x, pos, qe, mask, w, autocast=True)

    enc, dec = enc.to(cuda), dec.to(cuda)
    for p in list(enc.parameters()) + list(dec.parameters()):
        p.grad = None
    pos_d, mask_d = positional_encoding_tokens(eh, ew, heights.to(cuda), widths.to(cuda), 32, 128, 10000)
    assert (pos_d.cpu() - pos).abs().max() <= 1e-5 and torch.equal(mask_d.cpu(), mask)
    xg = x.to(cuda).requires_grad_(True)
    qe_p = qe.to(cuda).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        mem = enc(xg, position_embedding=pos_d, key_padding_mask=mask_d)
        out = dec(mem, position_embedding=pos_d, object_query_embedding=qe_p[None].expand(B, -1, -1), key_padding_mask=mask_d)
    assert out.shape == (B, 6, Q, 256) and out.dtype == torch.float32
    (out.float() * w.to(cuda)).sum().backward()

    def gate(name, got, ref, ref_bf16, eps=2e-3):
        got, scale = got.float().cpu(), ref.abs().max().item() + 1e-12
        if name.endswith("key_proj.bias"):
            # mathematically zero (a constant added to every key shifts all scores of a row equally: softmax-invariant); what is
            # left is rounding noise on both sides -- measure it on the scale of the sibling query bias gradient
            scale = gp_r[name[len("grad "):].replace("key_proj", "query_proj")].abs().max().item() + 1e-12
        err, err_ref = (got - ref).abs().max().item(), (ref_bf16.float() - ref).abs().max().item()
        assert err <= 2 * err_ref + eps * scale, f"{name}: err {err:.3e} > 2 * {err_ref:.3e} + {eps * scale:.3e} (scale {scale:.3e})"
        return err / scale, err_ref / scale

    gate("memory", mem, mem_r, mem_b)
    gate("decoded", out, out_r, out_b)
    gate("grad_x", xg.grad, gx_r, gx_b)
    worst = ("", 0.0, 0.0)
    for prefix, mod in (("enc.", enc), ("dec.", dec)):
        for n, p in mod.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), prefix + n
            rel, rel_ref = gate("grad " + prefix + n, p.grad, gp_r[prefix + n], gp_b[prefix + n])
            if rel > worst[1]:
                worst = (prefix + n, rel, rel_ref)
    print(f"worst parameter gradient: {worst[0]} rel err {worst[1]:.3e} (reference bf16 path: {worst[2]:.3e})")
    # the query embedding is a Parameter of DETR (detr/model.py:39,81): its gradient flows through the "+ embedding" prologue
    gate("grad query_embed", qe_p.grad, gp_r["query_embed"], gp_b["query_embed"])
