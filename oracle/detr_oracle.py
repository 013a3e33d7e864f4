"""oracle/detr_oracle.py -- TEST INFRASTRUCTURE ONLY (the checker, never the product).

Plain torch-fp32 / CPU restatement, written functionally over a `state_dict`, of the
reference's data-parallel hot path:

  a1  ScaledDotProductAttention.forward          detr/model.py:254-356
  a2  EncoderLayer / Encoder.forward             detr/model.py:220-225 / 206-209
  a3  DecoderLayer / Decoder.forward             detr/model.py:165-183 / 137-151
  a4  FFN.forward                                detr/model.py:395-424
  a5  PositionalEncoding                         detr/position_encoding.py:5-97
  a6  DETR.make_image_padding_mask               detr/model.py:96-114
  a7  HungarianMatcher.forward                   detr/matcher.py:40-99
  a8  box_iou / generalized_box_iou              detr/utils.py:57-97
  a9  linear_sum_assignment                      -> oracle/lsap.c
  a10 SetCriterion.forward                       detr/loss.py:198-231
  a11 loss_labels  a12 loss_cardinality  a13 loss_boxes   detr/loss.py:57-164
  a16 XYXY<->CXCYWH conversion                   torchvision transforms/v2/functional/_meta.py:158-194

Parity pinning: tests/golden/make_golden.py imports the REAL reference from /root/reference
in the authoring container, runs it and this file on the same seeded inputs, asserts agreement
and stores the reference's outputs as fixtures; tests/test_oracle_golden.py re-checks this file
against those fixtures everywhere (CPU, no reference needed).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Dropout is a no-op here (parity is defined with dropout off).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import lsap_oracle

Tensor = torch.Tensor


# --------------------------------------------------------------------------- boxes (a8, a16)
def xyxy_to_cxcywh(b: Tensor) -> Tensor:
    """w = x2-x1 ; cx = (2*x1 + w)/2  (op order of torchvision _meta.py:185-194)."""
    wh = b[..., 2:] - b[..., :2]
    c = (b[..., :2] * 2 + wh) / 2
    return torch.cat([c, wh], dim=-1)


def cxcywh_to_xyxy(b: Tensor) -> Tensor:
    """x1 = cx - w/2 ; x2 = w + x1  (op order of torchvision _meta.py:158-182)."""
    lo = b[..., :2] - b[..., 2:] / 2
    hi = b[..., 2:] + lo
    return torch.cat([lo, hi], dim=-1)


def pairwise_giou(a: Tensor, b: Tensor) -> Tensor:
    """(N,4),(M,4) xyxy -> (N,M).  No eps, clamp(min=0) on extents (detr/utils.py:57-97)."""
    if not (bool((a[:, 2:] >= a[:, :2]).all()) and bool((b[:, 2:] >= b[:, :2]).all())):
        raise AssertionError("degenerate box")  # detr/utils.py:87-88
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lo = torch.maximum(a[:, None, :2], b[None, :, :2])
    hi = torch.minimum(a[:, None, 2:], b[None, :, 2:])
    ext = (hi - lo).clamp(min=0)
    inter = ext[..., 0] * ext[..., 1]
    union = area_a[:, None] + area_b[None, :] - inter
    iou = inter / union
    lo_c = torch.minimum(a[:, None, :2], b[None, :, :2])
    hi_c = torch.maximum(a[:, None, 2:], b[None, :, 2:])
    ext_c = (hi_c - lo_c).clamp(min=0)
    hull = ext_c[..., 0] * ext_c[..., 1]
    return iou - (hull - union) / hull


# --------------------------------------------------------------------------- matcher (a7, a9)
def cost_matrix(logits: Tensor, boxes: Tensor, labels: Tensor, gt_xyxy: Tensor,
                w_class: float = 1.0, w_bbox: float = 1.0, w_giou: float = 1.0) -> Tensor:
    """One image: (Q,K) logits, (Q,4) cxcywh boxes, (M,) labels, (M,4) XYXY gt -> (Q,M) fp32.

    C = w_bbox*L1 + w_class*(-softmax[:,label]) + w_giou*(-GIoU), summed left to right
    (detr/matcher.py:66-93)."""
    prob = logits.softmax(-1)
    c_class = -prob[:, labels]
    gt_c = xyxy_to_cxcywh(gt_xyxy)
    c_bbox = (boxes[:, None, :] - gt_c[None, :, :]).abs().sum(-1)
    c_giou = -pairwise_giou(cxcywh_to_xyxy(boxes), gt_xyxy)
    return w_bbox * c_bbox + w_class * c_class + w_giou * c_giou


def hungarian_match(batch_logits: Tensor, batch_boxes: Tensor, gt_labels: Sequence[Tensor],
                    gt_boxes: Sequence[Tensor], w_class: float = 1.0, w_bbox: float = 1.0,
                    w_giou: float = 1.0, return_cost: bool = False):
    """List over images of (query_idx ascending, gt_idx), int64 CPU tensors (detr/matcher.py:40-99)."""
    if w_class == 0 and w_bbox == 0 and w_giou == 0:
        raise AssertionError("all costs can't be 0")  # detr/matcher.py:38
    out, costs = [], []
    with torch.no_grad():
        for lg, bx, lab, gt in zip(batch_logits, batch_boxes, gt_labels, gt_boxes):
            c = cost_matrix(lg.float(), bx.float(), lab, gt.float(), w_class, w_bbox, w_giou)
            r, k = lsap_oracle.linear_sum_assignment(c.cpu().numpy())
            out.append((torch.from_numpy(r), torch.from_numpy(k)))
            costs.append(c)
    return (out, costs) if return_cost else out


# --------------------------------------------------------------------------- criterion (a10-a15)
def giou_loss_pairs(src_xyxy: Tensor, tgt_xyxy: Tensor, eps: float = 1e-7) -> Tensor:
    """Elementwise GIoU loss of torchvision.ops.generalized_box_iou_loss (giou_loss.py:47-62,
    _utils.py:87-106): intersection zeroed unless strictly positive extents; eps in both denominators."""
    x1, y1, x2, y2 = src_xyxy.unbind(-1)
    x1g, y1g, x2g, y2g = tgt_xyxy.unbind(-1)
    ix1, iy1 = torch.maximum(x1, x1g), torch.maximum(y1, y1g)
    ix2, iy2 = torch.minimum(x2, x2g), torch.minimum(y2, y2g)
    ok = (iy2 > iy1) & (ix2 > ix1)
    inter = torch.where(ok, (ix2 - ix1) * (iy2 - iy1), torch.zeros_like(x1))
    union = (x2 - x1) * (y2 - y1) + (x2g - x1g) * (y2g - y1g) - inter
    iou = inter / (union + eps)
    hx1, hy1 = torch.minimum(x1, x1g), torch.minimum(y1, y1g)
    hx2, hy2 = torch.maximum(x2, x2g), torch.maximum(y2, y2g)
    hull = (hx2 - hx1) * (hy2 - hy1)
    return 1 - (iou - (hull - union) / (hull + eps))


def criterion_layer(logits: Tensor, boxes: Tensor, gt_labels: Sequence[Tensor], gt_boxes: Sequence[Tensor],
                    indices, num_classes: int, eos_coef: float = 0.1, w_ce: float = 1.0, w_l1: float = 5.0,
                    w_giou: float = 2.0) -> Dict[str, Tensor]:
    """One decoder layer: logits (B,Q,K), boxes (B,Q,4). Returns ce, cardinality, l1, giou, class_error."""
    B, Q, K = logits.shape
    dev = logits.device
    tgt = torch.full((B, Q), num_classes, dtype=torch.int64, device=dev)
    for b, (qi, gi) in enumerate(indices):
        tgt[b, qi.to(dev)] = gt_labels[b][gi.to(gt_labels[b].device)].to(dev)
    w = torch.ones(K, device=dev, dtype=logits.dtype)
    w[-1] = eos_coef
    # weighted mean: sum_i w[t_i]*nll_i / sum_i w[t_i]   (detr/loss.py:90)
    logp = F.log_softmax(logits.reshape(B * Q, K), dim=-1)
    t = tgt.reshape(-1)
    nll = -logp.gather(1, t[:, None])[:, 0]
    wt = w[t]
    ce = (wt * nll).sum() / wt.sum() * w_ce
    # cardinality (detr/loss.py:97-121), no grad
    with torch.no_grad():
        n_gt = torch.tensor([len(l) for l in gt_labels], dtype=torch.float32, device=dev)
        n_pred = (logits.argmax(-1) != K - 1).sum(1).float()
        card = (n_pred - n_gt).abs().mean()
        # class error on matched pairs (detr/utils.py:100-116); 100 when there are none
        sel_b = torch.cat([torch.full_like(qi, b) for b, (qi, _) in enumerate(indices)]) if B else torch.zeros(0, dtype=torch.int64)
        sel_q = torch.cat([qi for qi, _ in indices]) if B else torch.zeros(0, dtype=torch.int64)
        if sel_q.numel():
            hit = (logits[sel_b.to(dev), sel_q.to(dev)].argmax(-1) == tgt[sel_b.to(dev), sel_q.to(dev)]).float().sum()
            class_error = 100 - hit * (100.0 / sel_q.numel())
        else:
            class_error = torch.full((), 100.0, device=dev)
    # boxes (detr/loss.py:123-164): normaliser = local total GT count, clamped to 1
    n_norm = max(sum(len(g) for g in gt_boxes), 1)
    if sel_q.numel():
        src = boxes[sel_b.to(dev), sel_q.to(dev)]
        tg = torch.cat([g[gi.to(g.device)] for g, (_, gi) in zip(gt_boxes, indices)], dim=0).to(dev)
        l1 = (src - xyxy_to_cxcywh(tg)).abs().sum() * w_l1 / n_norm
        giou = giou_loss_pairs(cxcywh_to_xyxy(src), tg).sum() * w_giou / n_norm
    else:
        l1 = boxes.sum() * 0
        giou = boxes.sum() * 0
    return {"ce": ce, "card": card, "l1": l1, "giou": giou, "class_error": class_error}


def set_criterion(outputs: Dict[str, Tensor], targets: Dict[str, list], num_classes: int,
                  matcher_w=(1.0, 1.0, 1.0), eos_coef: float = 0.1, w_ce: float = 1.0, w_l1: float = 5.0,
                  w_giou: float = 2.0, return_indices: bool = False):
    """25-key dict of detr/loss.py:198-231 (matcher re-run for every decoder layer)."""
    lg_all, bx_all = outputs["pred_logits"], outputs["pred_boxes"]
    L = lg_all.shape[1]
    out: Dict[str, Tensor] = {}
    all_idx = []
    for l in range(L):
        lg, bx = lg_all[:, l], bx_all[:, l]
        idx = hungarian_match(lg.detach(), bx.detach(), targets["class_idx"], targets["boxes_normalized"], *matcher_w)
        all_idx.append(idx)
        r = criterion_layer(lg, bx, targets["class_idx"], targets["boxes_normalized"], idx, num_classes,
                            eos_coef, w_ce, w_l1, w_giou)
        sfx = f"_{l}" if l < L - 1 else ""
        if l == L - 1:
            out["class_error"] = r["class_error"]
        out[f"loss_label_ce{sfx}"] = r["ce"]
        out[f"cardinality_error{sfx}"] = r["card"]
        out[f"loss_l1_bbox{sfx}"] = r["l1"]
        out[f"loss_giou{sfx}"] = r["giou"]
    return (out, all_idx) if return_indices else out


# --------------------------------------------------------------------------- transformer (a1-a6)
def sdpa(sd: Dict[str, Tensor], prefix: str, query: Tensor, key: Tensor, value: Tensor, n_head: int,
         key_padding_mask: Optional[Tensor] = None, attention_mask: Optional[Tensor] = None) -> Tensor:
    """detr/model.py:254-356 with dropout off.  Masked scores become finfo.min (NOT -inf), so a fully
    masked row yields a uniform distribution."""
    B, Lq, C = query.shape
    S = key.shape[1]
    d = C // n_head
    q = F.linear(query, sd[prefix + "query_proj.weight"], sd[prefix + "query_proj.bias"])
    k = F.linear(key, sd[prefix + "key_proj.weight"], sd[prefix + "key_proj.bias"])
    v = F.linear(value, sd[prefix + "value_proj.weight"], sd[prefix + "value_proj.bias"])
    q = q.reshape(B, Lq, n_head, d).permute(0, 2, 1, 3)
    k = k.reshape(B, S, n_head, d).permute(0, 2, 3, 1)
    v = v.reshape(B, S, n_head, d).permute(0, 2, 1, 3)
    s = torch.matmul(q, k) / math.sqrt(d)
    neg = torch.finfo(s.dtype).min
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], neg)
    if attention_mask is not None:
        s = s.masked_fill(attention_mask, neg)
    p = s.softmax(-1)
    y = torch.matmul(p, v).permute(0, 2, 1, 3).reshape(B, Lq, C)
    return F.linear(y, sd[prefix + "output_proj.weight"], sd[prefix + "output_proj.bias"])


def _ln(sd, prefix, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + "weight"], sd[prefix + "bias"], eps)


def ffn(sd, prefix, x):
    """Linear -> GELU(tanh) -> Linear (detr/model.py:395-424); keys ffn.layers.0 / ffn.layers.3."""
    h = F.gelu(F.linear(x, sd[prefix + "layers.0.weight"], sd[prefix + "layers.0.bias"]), approximate="tanh")
    return F.linear(h, sd[prefix + "layers.3.weight"], sd[prefix + "layers.3.bias"])


def encoder(sd: Dict[str, Tensor], x: Tensor, pos: Tensor, key_padding_mask: Optional[Tensor], n_layers: int,
            n_head: int, eps: float = 1e-5) -> Tensor:
    """detr/model.py:206-209,220-225.  `sd` holds the Encoder's own state_dict (keys layers.N....)."""
    for n in range(n_layers):
        p = f"layers.{n}."
        a = _ln(sd, p + "norm1.", x, eps)
        qk = a + pos
        x = x + sdpa(sd, p + "self_attention.", qk, qk, a, n_head, key_padding_mask)
        x = x + ffn(sd, p + "ffn.", _ln(sd, p + "norm2.", x, eps))
    return _ln(sd, "norm.", x, eps)


def decoder(sd: Dict[str, Tensor], memory: Tensor, pos: Tensor, query_embed: Tensor,
            key_padding_mask: Optional[Tensor], n_layers: int, n_head: int, eps: float = 1e-5) -> Tensor:
    """detr/model.py:137-151,165-183 -> (B, n_layers, Q, C); the shared final LN is applied to every layer."""
    x = torch.zeros_like(query_embed)
    outs = []
    k_cross = memory + pos
    for n in range(n_layers):
        p = f"layers.{n}."
        a = _ln(sd, p + "norm1.", x, eps)
        qk = a + query_embed
        x = x + sdpa(sd, p + "self_attention.", qk, qk, a, n_head)
        a = _ln(sd, p + "norm2.", x, eps)
        x = x + sdpa(sd, p + "cross_attention.", a + query_embed, k_cross, memory, n_head, key_padding_mask)
        x = x + ffn(sd, p + "ffn.", _ln(sd, p + "norm3.", x, eps))
        outs.append(_ln(sd, "norm.", x, eps))
    return torch.stack(outs, dim=1)


def positional_encoding(embed_h: int, embed_w: int, heights: Tensor, widths: Tensor, scale: int = 32,
                        num_pos_feats: int = 128, temperature: float = 10000.0) -> Tensor:
    """detr/position_encoding.py:5-97 -> (B, 2*num_pos_feats, embed_h, embed_w) fp32.
    Coordinates are linspace(0,1,n) inside the valid ceil(h/scale) x ceil(w/scale) window, 0 in padding."""
    B = len(heights)
    hs = torch.ceil(heights / scale).to(torch.int64).tolist()
    ws = torch.ceil(widths / scale).to(torch.int64).tolist()
    gx = torch.zeros(B, embed_h, embed_w)
    gy = torch.zeros(B, embed_h, embed_w)
    for b in range(B):
        h, w = int(hs[b]), int(ws[b])
        gx[b, :h, :w] = torch.linspace(0, 1, w)[None, :].expand(h, w)
        gy[b, :h, :w] = torch.linspace(0, 1, h)[:, None].expand(h, w)
    gx, gy = gx * (2 * torch.pi), gy * (2 * torch.pi)
    freq = temperature ** (torch.arange(0, num_pos_feats, 2, dtype=torch.float32) / num_pos_feats)
    ax, ay = gx[..., None] / freq, gy[..., None] / freq
    px = torch.stack((ax.sin(), ax.cos()), dim=-1).flatten(-2)
    py = torch.stack((ay.sin(), ay.cos()), dim=-1).flatten(-2)
    return torch.cat((py, px), dim=-1).permute(0, 3, 1, 2).to(heights.device)


def padding_mask(embed_h: int, embed_w: int, heights: Tensor, widths: Tensor, scale: int = 32) -> Tensor:
    """detr/model.py:96-114: True ONLY on the bottom-right corner [ceil(h/s):, ceil(w/s):] (reference quirk)."""
    B = len(heights)
    m = torch.zeros(B, embed_h, embed_w, dtype=torch.bool)
    hs = torch.ceil(heights / scale).to(torch.int64).tolist()
    ws = torch.ceil(widths / scale).to(torch.int64).tolist()
    for b in range(B):
        m[b, int(hs[b]):, int(ws[b]):] = True
    return m.to(heights.device)


# --------------------------------------------------------------------------- synthetic inputs
def synth_targets(batch: int, max_gt: int, num_classes: int, seed: int, min_gt: int = 1):
    """GT lists as in SURVEY.md section 8(d): centres U(0.2,0.8), sizes U(0.02,0.32), XYXY, labels int64."""
    g = torch.Generator().manual_seed(seed)
    n = torch.randint(min_gt, max_gt + 1, (batch,), generator=g).tolist()
    labels, boxes = [], []
    for m in n:
        c = torch.rand(m, 2, generator=g) * 0.6 + 0.2
        s = torch.rand(m, 2, generator=g) * 0.30 + 0.02
        boxes.append(torch.cat([c - s / 2, c + s / 2], dim=1).float())
        labels.append(torch.randint(0, num_classes, (m,), generator=g, dtype=torch.int64))
    return labels, boxes


def synth_predictions(batch: int, layers: int, queries: int, num_classes: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(batch, layers, queries, num_classes + 1, generator=g)
    boxes = torch.randn(batch, layers, queries, 4, generator=g).sigmoid()
    return logits, boxes
