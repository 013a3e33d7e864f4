"""oracle/detr_oracle.py against the fixtures produced by the REAL reference (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import detr_oracle as O
from util import GOLDEN, golden_losses, golden_targets, load_golden

CRIT = ["criterion_q100", "criterion_q20_tall", "criterion_allempty"]


@pytest.mark.parametrize("name", CRIT)
def test_matcher_cost_and_indices(name):
    fx = load_golden(name)
    logits, boxes = torch.from_numpy(fx["logits"]), torch.from_numpy(fx["boxes"])
    tg = golden_targets(fx)
    w = fx["matcher_w"].tolist()
    B, L = logits.shape[:2]
    for l in range(L):
        idx, cost = O.hungarian_match(logits[:, l], boxes[:, l], tg["class_idx"], tg["boxes_normalized"], *w, return_cost=True)
        for b in range(B):
            np.testing.assert_allclose(cost[b].numpy(), fx[f"cost/{l}/{b}"], rtol=0, atol=2e-6)
            assert np.array_equal(idx[b][0].numpy(), fx[f"idx_q/{l}/{b}"])
            assert np.array_equal(idx[b][1].numpy(), fx[f"idx_gt/{l}/{b}"])


@pytest.mark.parametrize("name", CRIT)
def test_criterion_losses_and_grads(name):
    fx = load_golden(name)
    logits = torch.from_numpy(fx["logits"]).requires_grad_(True)
    boxes = torch.from_numpy(fx["boxes"]).requires_grad_(True)
    out = O.set_criterion({"pred_logits": logits, "pred_boxes": boxes}, golden_targets(fx), int(fx["num_classes"]),
                          tuple(fx["matcher_w"].tolist()))
    ref = golden_losses(fx)
    assert set(out) == set(ref)
    for k, v in ref.items():
        assert float(out[k]) == pytest.approx(v, rel=2e-6, abs=2e-6), k
    sum(v for k, v in out.items() if k.startswith("loss")).backward()
    np.testing.assert_allclose(logits.grad.numpy(), fx["grad_logits"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(boxes.grad.numpy(), fx["grad_boxes"], rtol=0, atol=1e-7)


def test_transformer_tiny():
    fx = load_golden("transformer_tiny")
    C, nh, _, ne, nd, Q = fx["cfg"].tolist()
    esd = {k[4:]: torch.from_numpy(v) for k, v in fx.items() if k.startswith("enc/")}
    dsd = {k[4:]: torch.from_numpy(v) for k, v in fx.items() if k.startswith("dec/")}
    x = torch.from_numpy(fx["x"]).requires_grad_(True)
    pos, mask = torch.from_numpy(fx["pos"]), torch.from_numpy(fx["mask"])
    qe = torch.from_numpy(fx["query_embed"])[None].repeat(x.shape[0], 1, 1)
    mem = O.encoder(esd, x, pos, mask, ne, nh)
    out = O.decoder(dsd, mem, pos, qe, mask, nd, nh)
    np.testing.assert_allclose(mem.detach().numpy(), fx["memory"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out.detach().numpy(), fx["decoded"], rtol=0, atol=2e-5)
    (out * torch.from_numpy(fx["w_out"])).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), fx["grad_x"], rtol=0, atol=1e-5)
    assert mask.any(), "fixture must exercise the padding mask"


def test_positional_encoding_and_mask():
    fx = load_golden("transformer_tiny")
    eh, ew = fx["embed_hw"].tolist()
    h, w = torch.from_numpy(fx["heights"]), torch.from_numpy(fx["widths"])
    pos = O.positional_encoding(eh, ew, h, w, 32, fx["pos_chw"].shape[1] // 2)
    np.testing.assert_allclose(pos.numpy(), fx["pos_chw"], rtol=0, atol=1e-6)
    m = O.padding_mask(eh, ew, h, w, 32)
    assert np.array_equal(m.flatten(1).numpy(), fx["mask"])


def test_pinning_report_is_green():
    rep = json.load(open(os.path.join(GOLDEN, "PINNING.json")))
    for k in CRIT:
        assert rep[k]["indices"] == "identical" and rep[k]["oracle_vs_ref_cost_maxabs"] < 2e-6
    assert rep["transformer_full_check"]["oracle_vs_ref_maxabs"]["out"] < 2e-5
