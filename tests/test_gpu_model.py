"""GPU parity of Encoder / Decoder (reference signatures + state_dict keys) against the golden fixture produced
by the real reference and against the fp32 oracle at full size.  Dropout off (eval), as parity is defined.

Tolerance: the attention core runs bf16 tensor-core operands with fp32 softmax/accumulation inside an otherwise
fp32 (or bf16-autocast) model; outputs are LayerNorm-ed, O(1).  Bar: max abs err <= 3e-2 and mean abs err <= 3e-3
for activations in fp32 mode; under bf16 autocast the error must not exceed 2x the reference-style bf16 autocast
oracle's own error + 1e-2."""
import numpy as np
import pytest
import torch

from oracle import detr_oracle as O
from util import load_golden

pytestmark = pytest.mark.gpu


def _cfg(C, nh, ffn, ne, nd, Q):
    from detr_b200.model import DETRConfig
    return DETRConfig(num_object_queries=Q, num_encoder_layers=ne, num_decoder_layers=nd, num_attention_heads=nh,
                      hidden_size=C, ffn_scale_factor=ffn)


def test_state_dict_keys_match_reference_layout():
    from detr_b200.model import Decoder, Encoder
    fx = load_golden("transformer_tiny")
    cfg = _cfg(*fx["cfg"].tolist())
    enc, dec = Encoder(cfg), Decoder(cfg)
    ref_e = {k[4:]: v.shape for k, v in fx.items() if k.startswith("enc/")}
    ref_d = {k[4:]: v.shape for k, v in fx.items() if k.startswith("dec/")}
    assert {k: tuple(v.shape) for k, v in enc.state_dict().items()} == ref_e
    assert {k: tuple(v.shape) for k, v in dec.state_dict().items()} == ref_d


def test_encoder_decoder_vs_golden(cuda):
    from detr_b200.model import Decoder, Encoder
    fx = load_golden("transformer_tiny")
    cfg = _cfg(*fx["cfg"].tolist())
    enc, dec = Encoder(cfg).to(cuda).eval(), Decoder(cfg).to(cuda).eval()
    enc.load_state_dict({k[4:]: torch.from_numpy(v) for k, v in fx.items() if k.startswith("enc/")})
    dec.load_state_dict({k[4:]: torch.from_numpy(v) for k, v in fx.items() if k.startswith("dec/")})
    x = torch.from_numpy(fx["x"]).to(cuda).requires_grad_(True)
    pos, mask = torch.from_numpy(fx["pos"]).to(cuda), torch.from_numpy(fx["mask"]).to(cuda)
    qe = torch.from_numpy(fx["query_embed"]).to(cuda)[None].expand(x.shape[0], -1, -1)
    mem = enc(x, position_embedding=pos, key_padding_mask=mask)
    out = dec(mem, position_embedding=pos, object_query_embedding=qe, key_padding_mask=mask)
    assert out.shape == fx["decoded"].shape
    e_mem = np.abs(mem.detach().cpu().numpy() - fx["memory"])
    e_out = np.abs(out.detach().cpu().numpy() - fx["decoded"])
    assert e_mem.max() <= 3e-2 and e_mem.mean() <= 3e-3, (e_mem.max(), e_mem.mean())
    assert e_out.max() <= 3e-2 and e_out.mean() <= 3e-3, (e_out.max(), e_out.mean())
    (out * torch.from_numpy(fx["w_out"]).to(cuda)).sum().backward()
    g, gr = x.grad.cpu().numpy(), fx["grad_x"]
    assert np.abs(g - gr).max() <= 5e-2 * np.abs(gr).max(), (np.abs(g - gr).max(), np.abs(gr).max())


@pytest.mark.parametrize("autocast", [False, True])
def test_full_size_encoder_decoder_vs_oracle(cuda, autocast):
    """Default DETRConfig (C=256, 8 heads, 6+6 layers, Q=100) on a 10x13 feature map with the corner mask."""
    from detr_b200.harness import padding_mask_device, positional_encoding_device
    from detr_b200.model import DETRConfig, Decoder, Encoder
    torch.manual_seed(3)
    cfg = DETRConfig(num_classes=91)
    enc, dec = Encoder(cfg).eval(), Decoder(cfg).eval()
    with torch.no_grad():
        for p in list(enc.parameters()) + list(dec.parameters()):
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    B, eh, ew = 2, 10, 13
    heights, widths = torch.tensor([320, 200], dtype=torch.int32), torch.tensor([416, 300], dtype=torch.int32)
    x = torch.randn(B, eh * ew, 256)
    qe = 0.5 * torch.randn(100, 256)
    pos = O.positional_encoding(eh, ew, heights, widths).flatten(2).permute(0, 2, 1)
    mask = O.padding_mask(eh, ew, heights, widths).flatten(1)
    # vectorised device versions of the two host loops
    pos_d = positional_encoding_device(eh, ew, heights.to(cuda), widths.to(cuda)).flatten(2).permute(0, 2, 1)
    assert (pos_d.cpu() - pos).abs().max() <= 1e-5
    assert torch.equal(padding_mask_device(eh, ew, heights.to(cuda), widths.to(cuda)).flatten(1).cpu(), mask)

    xr = x.clone().requires_grad_(True)
    esd, dsd = dict(enc.named_parameters()), dict(dec.named_parameters())
    mem_r = O.encoder(esd, xr, pos, mask, 6, 8)
    out_r = O.decoder(dsd, mem_r, pos, qe[None].expand(B, -1, -1), mask, 6, 8)
    w = torch.randn_like(out_r)
    (out_r * w).sum().backward()
    # the reference's own bf16-autocast error, for scale (CPU autocast of the oracle)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out_b = O.decoder(dsd, O.encoder(esd, x, pos, mask, 6, 8), pos, qe[None].expand(B, -1, -1), mask, 6, 8)
    err_ref = (out_b.float() - out_r).abs().max().item()

    enc, dec = enc.to(cuda), dec.to(cuda)
    xg = x.to(cuda).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        mem = enc(xg, position_embedding=pos_d, key_padding_mask=mask.to(cuda))
        out = dec(mem, position_embedding=pos_d, object_query_embedding=qe.to(cuda)[None].expand(B, -1, -1),
                  key_padding_mask=mask.to(cuda))
    assert out.shape == (B, 6, 100, 256)
    err = (out.float().cpu() - out_r).abs().max().item()
    if autocast:
        assert err <= 2 * err_ref + 1e-2, (err, err_ref)
    else:
        assert err <= 3e-2 and err <= 2 * err_ref + 1e-3, (err, err_ref)
    (out.float() * w.to(cuda)).sum().backward()
    gscale = xr.grad.abs().max().item()
    gerr = (xg.grad.cpu() - xr.grad).abs().max().item()
    assert gerr <= (0.15 if autocast else 0.06) * gscale, (gerr, gscale)
    # parameter gradients reach every tensor
    for n, p in list(enc.named_parameters()) + list(dec.named_parameters()):
        assert p.grad is not None and torch.isfinite(p.grad).all(), n


def test_train_mode_dropout_runs_and_is_seeded(cuda):
    from detr_b200.model import DETRConfig, Encoder
    cfg = DETRConfig()
    enc = Encoder(cfg).to(cuda).train()
    x = torch.randn(2, 130, 256, device=cuda)
    pos = torch.randn(2, 130, 256, device=cuda)
    mask = torch.zeros(2, 130, dtype=torch.bool, device=cuda)
    torch.manual_seed(5); a = enc(x, pos, mask)
    torch.manual_seed(5); b = enc(x, pos, mask)
    torch.manual_seed(6); c = enc(x, pos, mask)
    assert torch.equal(a, b) and not torch.equal(a, c) and torch.isfinite(a).all()


@pytest.mark.parametrize("M,N", [(6800, 2048), (6800, 256), (800, 512), (37, 8), (1, 256), (5000, 264)])
def test_colsum_kernel(cuda, M, N):
    from detr_b200.rowops import colsum
    g = torch.randn(M, N, device=cuda).bfloat16()
    ref = g.double().sum(0)
    got = colsum(g)
    assert got.dtype == torch.float32
    assert (got.double() - ref).abs().max().item() <= 1e-3 * max(1.0, M ** 0.5)
    v = g[:, : N // 2] if (N // 2) % 8 == 0 else g
    assert (colsum(v).double() - v.double().sum(0)).abs().max().item() <= 1e-3 * max(1.0, M ** 0.5)


def test_linear_fn_matches_autocast_linear(cuda):
    from detr_b200.rowops import linear
    torch.manual_seed(0)
    lin = torch.nn.Linear(256, 512).to(cuda)
    x = torch.randn(4, 333, 256, device=cuda, requires_grad=True)
    w = torch.randn(4, 333, 512, device=cuda)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y1 = linear(x, lin.weight, lin.bias)
    (y1.float() * w).sum().backward()
    g1 = (x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    x.grad = None; lin.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y2 = torch.nn.functional.linear(x, lin.weight, lin.bias)
    (y2.float() * w).sum().backward()
    assert torch.equal(y1, y2)
    assert torch.allclose(g1[0], x.grad, rtol=1e-2, atol=1e-2) and g1[1].dtype == torch.float32
    assert torch.allclose(g1[1], lin.weight.grad, rtol=2e-2, atol=2e-1)
    ref_b = w.sum((0, 1))
    assert (g1[2] - ref_b).abs().max() <= (lin.bias.grad - ref_b).abs().max() + 0.5   # at least as accurate as ATen's bf16 reduction


@pytest.mark.parametrize("xdt,odt,C", [(torch.float32, torch.float32, 256), (torch.bfloat16, torch.bfloat16, 256),
                                       (torch.float32, torch.bfloat16, 256), (torch.float32, torch.float32, 64), (torch.bfloat16, torch.bfloat16, 512)])
def test_fused_layernorm_add(cuda, xdt, odt, C):
    """Fused LN(+addend) forward/backward against F.layer_norm autograd in fp32."""
    from detr_b200.rowops import _LayerNormAdd
    torch.manual_seed(1)
    B, R = 3, 157
    x = torch.randn(B, R, C, device=cuda).mul(2).add(0.5).to(xdt)
    ln = torch.nn.LayerNorm(C).to(cuda)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5); ln.bias.normal_(0, 0.2)
    emb = torch.randn(R, C, device=cuda, requires_grad=True)
    w1, w2 = torch.randn(B, R, C, device=cuda), torch.randn(B, R, C, device=cuda)
    xa = x.clone().requires_grad_(True)
    y, y2 = _LayerNormAdd.apply(xa, ln.weight, ln.bias, emb[None].expand(B, -1, -1), ln.eps, odt, True)
    assert y.dtype == odt and y2.dtype == odt
    (y.float() * w1 + y2.float() * w2).sum().backward()
    got = (y.float(), y2.float(), xa.grad.float(), ln.weight.grad.clone(), ln.bias.grad.clone(), emb.grad.clone())
    ln.zero_grad(); emb.grad = None
    xr = x.float().clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (C,), ln.weight, ln.bias, ln.eps)
    y2r = yr + emb[None]
    (yr * w1 + y2r * w2).sum().backward()
    ref = (yr, y2r, xr.grad, ln.weight.grad, ln.bias.grad, emb.grad)
    lo = xdt == torch.bfloat16 or odt == torch.bfloat16
    for name, a, b in zip(("y", "y2", "dx", "dgamma", "dbeta", "dadd"), got, ref):
        tol = (3e-2 if lo else 2e-5) * max(1.0, b.abs().max().item())
        assert (a - b).abs().max().item() <= tol, (name, (a - b).abs().max().item(), tol)
    # y2-only variant (decoder cross-attention query)
    _, q = _LayerNormAdd.apply(x, ln.weight, ln.bias, emb[None].expand(B, -1, -1), ln.eps, odt, False)
    assert _ is None and (q.float() - y2r).abs().max().item() <= (3e-2 if lo else 2e-5) * y2r.abs().max().item()


@pytest.mark.parametrize("mode,res_dtype", [(0, torch.float32), (0, torch.bfloat16), (1, None)])
def test_fused_linear_epilogue_matches_torch(cuda, mode, res_dtype):
    """residual + dropout(lin(x)) and dropout(gelu_tanh(lin(x))) with dropout OFF against the unfused autocast path
    (detr/model.py:223-224, 405-411): values, dx, dW, db and the residual gradient."""
    from detr_b200.rowops import linear_dropout_add, linear_gelu_dropout
    torch.manual_seed(0)
    K, N = (256, 256) if mode == 0 else (256, 2048)
    lin = torch.nn.Linear(K, N).to(cuda)
    x = torch.randn(3, 211, K, device=cuda).bfloat16().requires_grad_(True)
    res = torch.randn(3, 211, N, device=cuda).to(res_dtype).requires_grad_(True) if mode == 0 else None
    w = torch.randn(3, 211, N, device=cuda)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = linear_dropout_add(x, res, lin, 0.0) if mode == 0 else linear_gelu_dropout(x, lin, 0.0)
    (out.float() * w).sum().backward()
    got = (out, x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone(), res.grad.clone() if mode == 0 else None)
    x.grad = None; lin.zero_grad()
    if mode == 0:
        res.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = torch.nn.functional.linear(x, lin.weight, lin.bias)
        ref = res + y if mode == 0 else torch.nn.functional.gelu(y, approximate="tanh")
    (ref.float() * w).sum().backward()
    assert out.dtype == ref.dtype and out.shape == ref.shape
    assert torch.allclose(out.float(), ref.float(), rtol=2e-2, atol=2e-2)
    assert torch.allclose(got[1].float(), x.grad.float(), rtol=3e-2, atol=3e-2)
    assert got[2].dtype == torch.float32 and torch.allclose(got[2], lin.weight.grad, rtol=3e-2, atol=3e-1)
    ref_b = lin.bias.grad
    assert torch.allclose(got[3], ref_b, rtol=3e-2, atol=0.5), (got[3] - ref_b).abs().max()
    if mode == 0:
        assert torch.allclose(got[4].float(), res.grad.float(), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("mode", [0, 1])
def test_fused_linear_epilogue_dropout_consistency(cuda, mode):
    """With dropout the backward must regenerate the forward's mask: dropped outputs get exactly zero gradient, kept
    ones the 1/keep-scaled gradient; the keep rate is 1 - round(128 p)/128; same seed -> same mask."""
    from detr_b200 import _lib
    M, N, p = 999, 256, 0.25
    y = (torch.randn(M, N, device=cuda) + 3.0).bfloat16()          # gelu(y) != 0 almost surely
    x = torch.zeros(M, N, device=cuda)
    outs = []
    for seed in (7, 7, 8):
        o = torch.empty(M, N, device=cuda, dtype=torch.float32 if mode == 0 else torch.bfloat16)
        _lib.call("detr_epilogue_fwd", mode, x.data_ptr() if mode == 0 else None, 0, y.data_ptr(), o.data_ptr(), M, N, p, seed, None, _lib.stream_ptr())
        outs.append(o.float())
    assert torch.equal(outs[0], outs[1]) and not torch.equal(outs[0], outs[2])
    kept = outs[0] != 0
    assert abs(kept.float().mean().item() - 0.75) < 0.01
    g = torch.ones(M, N, device=cuda)
    dy = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    chunks = _lib.load().detr_epilogue_chunks(M, N)
    partial = torch.empty(chunks * N, device=cuda)
    db = torch.empty(N, device=cuda)
    _lib.call("detr_epilogue_bwd", mode, g.data_ptr(), 0, y.data_ptr(), dy.data_ptr(), partial.data_ptr(), db.data_ptr(),
              _lib.zero_counters(g.device).data_ptr(), M, N, p, 7, None, _lib.stream_ptr())
    assert torch.equal(dy.float() != 0, kept)
    if mode == 0:
        assert torch.allclose(dy.float()[kept], torch.full((), 1 / 0.75, device=cuda).expand(int(kept.sum())), rtol=1e-2)
        assert torch.allclose(outs[0][kept], (y.float() / 0.75)[kept], rtol=1e-2)
    assert torch.allclose(db, dy.float().sum(0), rtol=1e-3, atol=1e-2)


def test_backbone_fold_pack_matches_per_conv_fold(cuda):
    """One-launch BN fold of all convolution weights (detr_scale_cast_multi) against folding each weight where it is
    used: same activations, same parameter gradients (both paths feed cuDNN the same bf16 weights)."""
    from detr_b200.harness import _Backbone
    torch.manual_seed(0)
    bb = _Backbone("resnet50").to(cuda).to(memory_format=torch.channels_last).train()
    for m in bb.modules():   # non-trivial frozen statistics
        if hasattr(m, "running_var"):
            m.running_var.uniform_(0.5, 1.5); m.running_mean.normal_(0, 0.1); m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.1)
    x = torch.randn(2, 3, 96, 128, device=cuda)
    res = []
    for use in (True, False):
        bb.use_fold_pack = use
        bb.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = bb(x)
        y.float().square().mean().backward()
        gs = {n: p.grad.clone() for n, p in bb.named_parameters() if p.grad is not None}
        res.append((y.float(), gs))
    (y1, g1), (y0, g0) = res
    with torch.no_grad():
        y_ref = bb(x)                       # fp32, no autocast: plain folded convolutions
    top = y_ref.abs().max().item()
    # bf16 activations through 53 convolutions: a few percent of the output range, for either path against fp32 and between them
    assert (y1 - y_ref).abs().max().item() <= 5e-2 * top and (y0 - y_ref).abs().max().item() <= 5e-2 * top
    assert (y1 - y0).abs().max().item() <= 5e-2 * top
    assert set(g1) == set(g0) and len(g1) >= 53
    for n in g0:
        assert g1[n].dtype == g0[n].dtype and g1[n].shape == g0[n].shape
        # two bf16 executions of a 53-convolution random network drift apart towards the stem (ReLU masks flip on rounding
        # differences): relative L2 error, tight next to the loss, looser at the far end
        rel = ((g1[n] - g0[n]).norm() / (g0[n].norm() + 1e-12)).item()
        assert rel <= (5e-2 if "layer4" in n else 2e-1), (n, rel)


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 50, 68), (1, 8, 7, 9), (2, 64, 33, 31)])
def test_stem_maxpool_kernels_match_aten(cuda, B, C, H, W):
    """3x3 / stride 2 / pad 1 channels_last bf16 max pooling: values bit-exact, gradients equal to ATen's (including its
    first-maximum tie rule: the input is ReLU-like with many exact ties)."""
    from detr_b200.harness import _stem_maxpool
    torch.manual_seed(0)
    pool = torch.nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
    x0 = torch.randn(B, C, H, W, device=cuda).relu().bfloat16().contiguous(memory_format=torch.channels_last)
    xa, xb = x0.clone().requires_grad_(True), x0.clone().requires_grad_(True)
    ya, yb = _stem_maxpool(pool, xa), pool(xb)
    assert ya.shape == yb.shape and torch.equal(ya, yb)
    w = torch.randn_like(yb)
    (ya * w).sum().backward(); (yb * w).sum().backward()
    assert torch.allclose(xa.grad.float(), xb.grad.float(), rtol=1e-2, atol=1e-2)
    assert ((xa.grad != 0) != (xb.grad != 0)).float().mean().item() < 1e-4   # same arg-max choice everywhere (ties included)


def test_backbone_block_backward_fusion_matches_unfused(cuda):
    """One autograd node per bottleneck with (dx + g_identity) * (x > 0) fused (detr_add_relu_mask_bf16) against the per-convolution
    autograd path: same activations (identical cuDNN calls), parameter gradients equal up to bf16 accumulation order."""
    from detr_b200.harness import _Backbone
    torch.manual_seed(0)
    bb = _Backbone("resnet50").to(cuda).to(memory_format=torch.channels_last).train()
    x = torch.randn(2, 3, 96, 128, device=cuda)
    res = []
    for fuse in (True, False):
        bb.fuse_block_backward = fuse
        bb.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = bb(x)
        y.float().square().mean().backward()
        res.append((y.float(), {n: p.grad.clone() for n, p in bb.named_parameters() if p.grad is not None}))
    (y1, g1), (y0, g0) = res
    assert torch.equal(y1, y0)
    assert set(g1) == set(g0)
    for n in g0:
        rel = ((g1[n] - g0[n]).norm() / (g0[n].norm() + 1e-12)).item()
        assert rel <= 2e-2, (n, rel)


def test_stem_space_to_depth_matches_direct_conv(cuda):
    """The 4x4 stride-1 convolution over the space-to-depth image equals the 7x7 stride-2 stem (same bf16 products, different
    summation order), forward and weight gradient."""
    from detr_b200.harness import _Backbone
    torch.manual_seed(0)
    bb = _Backbone("resnet50").to(cuda).to(memory_format=torch.channels_last).train()
    x = torch.randn(2, 3, 64, 96, device=cuda)
    res = []
    for s2d in (True, False):
        bb.stem_space_to_depth = s2d
        bb.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = bb(x)
        y.float().square().mean().backward()
        res.append((y.float(), bb.backbone.conv1.weight.grad.clone()))
    (y1, g1), (y0, g0) = res
    assert (y1 - y0).abs().max().item() <= 3e-2 * y0.abs().max().item()
    assert ((g1 - g0).norm() / g0.norm()).item() <= 5e-2


@pytest.mark.gpu
@pytest.mark.parametrize("xdt", [torch.float32, torch.bfloat16])
def test_fused_layernorm_passes_residual_gradient(cuda, xdt):
    """pass_x: x comes back as a third output (the block's residual input) and the gradient that returns through it is
    added to dx inside the LayerNorm backward kernel: x + f(LN(x)) against torch autograd in fp32."""
    from detr_b200.rowops import layer_norm_add
    torch.manual_seed(3)
    B, R, C = 2, 131, 256
    x = torch.randn(B, R, C, device=cuda).mul(1.5).to(xdt)
    ln = torch.nn.LayerNorm(C).to(cuda)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5); ln.bias.normal_(0, 0.2)
    w1, w2 = torch.randn(B, R, C, device=cuda), torch.randn(B, R, C, device=cuda)
    xa = x.clone().requires_grad_(True)
    y, y2, xres = layer_norm_add(xa, ln, None, pass_x=True)
    assert y2 is None and xres.shape == xa.shape and xres.dtype == xa.dtype
    out = xres.float() * w2 + y.float() * w1          # residual branch + f(LN(x))
    out.sum().backward()
    got = (xa.grad.float().clone(), ln.weight.grad.clone(), ln.bias.grad.clone())
    ln.zero_grad()
    xr = x.float().clone().requires_grad_(True)
    (xr * w2 + torch.nn.functional.layer_norm(xr, (C,), ln.weight, ln.bias, ln.eps) * w1).sum().backward()
    ref = (xr.grad, ln.weight.grad, ln.bias.grad)
    for name, a, b in zip(("dx", "dgamma", "dbeta"), got, ref):
        tol = (3e-2 if xdt == torch.bfloat16 else 2e-5) * max(1.0, b.abs().max().item())
        assert (a - b).abs().max().item() <= tol, (name, (a - b).abs().max().item(), tol)
    # only the residual branch is used: the gradient passes straight through
    xb = x.clone().requires_grad_(True)
    _, _, xres = layer_norm_add(xb, ln, None, pass_x=True)
    (xres.float() * w2).sum().backward()
    assert (xb.grad.float() - w2).abs().max().item() <= (2e-2 if xdt == torch.bfloat16 else 0.0) * w2.abs().max().item()
