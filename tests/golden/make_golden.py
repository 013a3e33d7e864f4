"""Generates tests/golden/*.npz by running the REAL reference (/root/reference, read-only) on seeded
synthetic inputs, and checks oracle/detr_oracle.py against it while doing so.

Run once in the authoring container:   python tests/golden/make_golden.py
The reference cannot travel to the GPU box, so its outputs are committed as fixtures; the tests only
read the fixtures.  Shims (both off the hot path): a dummy `torchmetrics` module
(detr/utils.py:3 imports it) -- nothing else is patched.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

tm = types.ModuleType("torchmetrics")
tmd = types.ModuleType("torchmetrics.detection")
tmd.MeanAveragePrecision = object
tm.detection = tmd
sys.modules["torchmetrics"] = tm
sys.modules["torchmetrics.detection"] = tmd

from detr.loss import SetCriterion  # noqa: E402  (reference)
from detr.matcher import HungarianMatcher  # noqa: E402
from detr.model import DETR, Decoder, DETRConfig, Encoder  # noqa: E402
from detr.position_encoding import PositionalEncoding  # noqa: E402

from oracle import detr_oracle as O  # noqa: E402

report = {}


def np_(t):
    return t.detach().cpu().numpy()


def maxerr(a, b):
    return float((a - b).abs().max()) if a.numel() else 0.0


# ----------------------------------------------------------------------------- matcher + criterion
def gen_criterion(name, B, L, Q, num_classes, gt_counts, seed):
    torch.manual_seed(seed)
    K = num_classes + 1
    logits = torch.randn(B, L, Q, K)
    boxes = torch.randn(B, L, Q, 4).sigmoid()
    labels, gts = [], []
    for m in gt_counts:
        c = torch.rand(m, 2) * 0.6 + 0.2
        s = torch.rand(m, 2) * 0.30 + 0.02
        gts.append(torch.cat([c - s / 2, c + s / 2], 1))
        labels.append(torch.randint(0, num_classes, (m,), dtype=torch.int64))
    targets = {"class_idx": labels, "boxes_normalized": gts}

    matcher = HungarianMatcher(cost_class=1.0, cost_bbox=5.0, cost_giou=2.0)  # detr/train.py:92-96
    crit = SetCriterion(num_classes, matcher, 1.0, 5.0, 2.0, 0.1)
    lg = logits.clone().requires_grad_(True)
    bx = boxes.clone().requires_grad_(True)
    ref = crit({"pred_logits": lg, "pred_boxes": bx}, targets)
    total = sum(v for k, v in ref.items() if k.startswith("loss"))
    total.backward()

    # reference cost matrices + indices, layer by layer (same code path as detr/matcher.py:66-97)
    fx = {"logits": np_(logits), "boxes": np_(boxes), "gt_counts": np.array(gt_counts, np.int64),
          "gt_labels": np_(torch.cat(labels)) if sum(gt_counts) else np.zeros(0, np.int64),
          "gt_boxes": np_(torch.cat(gts)) if sum(gt_counts) else np.zeros((0, 4), np.float32),
          "num_classes": np.int64(num_classes), "matcher_w": np.array([1.0, 5.0, 2.0]),
          "grad_logits": np_(lg.grad), "grad_boxes": np_(bx.grad)}
    for k, v in ref.items():
        fx["loss/" + k] = np_(v)
    from detr.utils import generalized_box_iou
    from torchvision.transforms.v2.functional import convert_bounding_box_format
    from torchvision.tv_tensors import BoundingBoxFormat as F_
    worst_cost = 0.0
    for l in range(L):
        idx = matcher(logits[:, l], boxes[:, l], labels, gts)
        probs = logits[:, l].softmax(-1)
        oidx, ocost = O.hungarian_match(logits[:, l], boxes[:, l], labels, gts, 1.0, 5.0, 2.0, return_cost=True)
        for b in range(B):
            cc = -probs[b][:, labels[b]]
            cb = torch.cdist(boxes[b, l], convert_bounding_box_format(gts[b], F_.XYXY, F_.CXCYWH), p=1)
            cg = -generalized_box_iou(convert_bounding_box_format(boxes[b, l], F_.CXCYWH, F_.XYXY), gts[b])
            C = 5.0 * cb + 1.0 * cc + 2.0 * cg
            fx[f"cost/{l}/{b}"] = np_(C)
            fx[f"idx_q/{l}/{b}"] = np_(idx[b][0])
            fx[f"idx_gt/{l}/{b}"] = np_(idx[b][1])
            worst_cost = max(worst_cost, maxerr(C, ocost[b]))
            assert torch.equal(idx[b][0], oidx[b][0]) and torch.equal(idx[b][1], oidx[b][1]), (name, l, b)
    # oracle criterion vs reference
    lg2 = logits.clone().requires_grad_(True)
    bx2 = boxes.clone().requires_grad_(True)
    mine = O.set_criterion({"pred_logits": lg2, "pred_boxes": bx2}, targets, num_classes, (1.0, 5.0, 2.0))
    sum(v for k, v in mine.items() if k.startswith("loss")).backward()
    assert set(mine) == set(ref), (set(mine) ^ set(ref))
    worst_loss = max(abs(float(mine[k]) - float(ref[k])) / max(1.0, abs(float(ref[k]))) for k in ref)
    g_err = max(maxerr(lg.grad, lg2.grad), maxerr(bx.grad, bx2.grad))
    assert worst_cost < 2e-6 and worst_loss < 2e-6 and g_err < 1e-7, (worst_cost, worst_loss, g_err)
    report[name] = {"oracle_vs_ref_cost_maxabs": worst_cost, "oracle_vs_ref_loss_rel": worst_loss,
                    "oracle_vs_ref_grad_maxabs": g_err, "indices": "identical"}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)


# ----------------------------------------------------------------------------- transformer
def gen_transformer(name, cfg, B, eh, ew, heights, widths, seed, store_weights):
    torch.manual_seed(seed)
    enc, dec = Encoder(cfg).eval(), Decoder(cfg).eval()
    # non-trivial LN affine + biases so that every parameter matters
    with torch.no_grad():
        for m in list(enc.parameters()) + list(dec.parameters()):
            if m.dim() == 1:
                m.add_(0.1 * torch.randn_like(m))
    C, Q = cfg.hidden_size, cfg.num_object_queries
    heights = torch.tensor(heights, dtype=torch.int32)
    widths = torch.tensor(widths, dtype=torch.int32)
    pos = PositionalEncoding(C // 2, cfg.temperature)(eh, ew, heights, widths, 32)
    mask = DETR.make_image_padding_mask(None, eh, ew, heights, widths, 32)
    opos = O.positional_encoding(eh, ew, heights, widths, 32, C // 2, cfg.temperature)
    omask = O.padding_mask(eh, ew, heights, widths, 32)
    assert torch.equal(mask, omask) and maxerr(pos, opos) < 1e-6, maxerr(pos, opos)
    x = torch.randn(B, eh * ew, C, requires_grad=True)
    posf = pos.flatten(2).permute(0, 2, 1)
    maskf = mask.flatten(1)
    qe = (0.5 * torch.randn(Q, C)).unsqueeze(0).repeat(B, 1, 1)
    mem = enc(x, position_embedding=posf, key_padding_mask=maskf)
    out = dec(mem, position_embedding=posf, object_query_embedding=qe, key_padding_mask=maskf)
    w_out = torch.randn_like(out)
    (out * w_out).sum().backward()

    x2 = x.detach().clone().requires_grad_(True)
    esd = {k: v.detach() for k, v in enc.state_dict().items()}
    dsd = {k: v.detach() for k, v in dec.state_dict().items()}
    omem = O.encoder(esd, x2, posf, maskf, cfg.num_encoder_layers, cfg.num_attention_heads, cfg.layer_norm_eps)
    oout = O.decoder(dsd, omem, posf, qe, maskf, cfg.num_decoder_layers, cfg.num_attention_heads, cfg.layer_norm_eps)
    (oout * w_out).sum().backward()
    e = {"mem": maxerr(mem, omem), "out": maxerr(out, oout), "dx": maxerr(x.grad, x2.grad)}
    assert e["mem"] < 2e-5 and e["out"] < 2e-5 and e["dx"] < 2e-4 * float(x.grad.abs().max()), e
    report[name] = {"oracle_vs_ref_maxabs": e}
    if store_weights:
        fx = {"x": np_(x), "pos": np_(posf), "mask": np_(maskf), "query_embed": np_(qe[0]), "w_out": np_(w_out),
              "memory": np_(mem), "decoded": np_(out), "grad_x": np_(x.grad), "pos_chw": np_(pos),
              "heights": np_(heights), "widths": np_(widths), "embed_hw": np.array([eh, ew]),
              "cfg": np.array([C, cfg.num_attention_heads, cfg.ffn_scale_factor, cfg.num_encoder_layers,
                               cfg.num_decoder_layers, Q])}
        for k, v in esd.items():
            fx["enc/" + k] = np_(v)
        for k, v in dsd.items():
            fx["dec/" + k] = np_(v)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)


if __name__ == "__main__":
    torch.set_num_threads(8)
    # matcher + criterion: empty image, M=1, typical, M==Q; then M>Q and Q=300
    gen_criterion("criterion_q100", B=4, L=2, Q=100, num_classes=91, gt_counts=[0, 1, 17, 100], seed=11)
    gen_criterion("criterion_q20_tall", B=3, L=3, Q=20, num_classes=7, gt_counts=[30, 20, 5], seed=12)
    gen_criterion("criterion_allempty", B=2, L=1, Q=10, num_classes=5, gt_counts=[0, 0], seed=13)
    # transformer: tiny config with stored weights (mask active: image 1 is smaller than the batch canvas)
    tiny = DETRConfig(num_object_queries=12, num_encoder_layers=2, num_decoder_layers=2, num_attention_heads=2,
                      hidden_size=64, ffn_scale_factor=2)
    gen_transformer("transformer_tiny", tiny, B=2, eh=5, ew=7, heights=[160, 97], widths=[224, 130], seed=21,
                    store_weights=True)
    # full-size config: checked here (oracle == reference), too large to store
    full = DETRConfig(num_classes=91)
    gen_transformer("transformer_full_check", full, B=2, eh=6, ew=8, heights=[192, 150], widths=[256, 200], seed=22,
                    store_weights=False)
    report["versions"] = {"torch": torch.__version__, "numpy": np.__version__}
    with open(os.path.join(HERE, "PINNING.json"), "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
    print(json.dumps(report, indent=1, sort_keys=True))
