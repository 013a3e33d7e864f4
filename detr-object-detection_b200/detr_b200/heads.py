"""Prediction heads on the tcgen05 GEMM kernels (detr/model.py:45-52,92-93,359-392):

    pred_logits = class_embedding(decoded)                 nn.Linear(256, num_classes + 1)
    pred_boxes  = bbox_embedding(decoded).sigmoid()        Linear -> GELU(tanh) -> Linear -> GELU(tanh) -> Linear(256, 4)

Forward: four launches of `detr_gemm_bf16` that write fp32 `(B, L, Q, num_classes + 1)` logits and fp32 sigmoid boxes
`(B, L, Q, 4)`, dense and 16-byte aligned -- exactly what `HungarianMatcher.match_layers` and the criterion kernels read, so
`SetCriterion`'s `.float()` / `_rows()` copy nothing.  Backward: one pass (`detr_heads_grad_prep`) turns the criterion's fp32
gradients into tile-wide bf16 operands (sigmoid backward included), then the input-gradient GEMMs (GELU backward in their
epilogues, the two branches joined by the residual epilogue) and the weight + bias gradient GEMMs.

The modules keep the reference's parameters and state_dict keys (`class_embedding.*`, `bbox_embedding.net.{0,2,4}.*`); this file
only replaces how they are applied.  Anything outside the kernels' contract (hidden size != 256, a box MLP that is not three
layers, more than 128 classes, no bf16 autocast) runs the plain composition."""
from __future__ import annotations

import weakref

import torch
from torch import nn

from . import _lib, gemm as G
from .rowops import ShadowedLinears

C_MODEL = 256
_CLS_PAD, _BOX_PAD = 128, 64      # rows the 92- and 4-row weights occupy in their shadows (whole MMA tiles for the backward GEMMs)
_SHADOWS: "weakref.WeakKeyDictionary[nn.Module, ShadowedLinears]" = weakref.WeakKeyDictionary()


def _mlp_linears(bbox_embedding: nn.Module):
    net = getattr(bbox_embedding, "net", None)
    if net is None:
        return None
    lins = [m for m in net if isinstance(m, nn.Linear)]
    acts = [m for m in net if not isinstance(m, nn.Linear)]
    ok = len(lins) == 3 and all(isinstance(a, nn.GELU) and a.approximate == "tanh" for a in acts) and len(acts) == 2
    return lins if ok else None


def supported(x: torch.Tensor, class_embedding: nn.Linear, bbox_embedding: nn.Module) -> bool:
    lins = _mlp_linears(bbox_embedding)
    if lins is None or not x.is_cuda or x.dim() < 2 or x.shape[-1] != C_MODEL:
        return False
    if not (torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16):
        return False
    n_cls = class_embedding.out_features
    shapes_ok = (class_embedding.in_features == C_MODEL and n_cls % 4 == 0 and 4 <= n_cls <= _CLS_PAD and class_embedding.bias is not None
                 and tuple(lins[0].weight.shape) == (C_MODEL, C_MODEL) and tuple(lins[1].weight.shape) == (C_MODEL, C_MODEL)
                 and tuple(lins[2].weight.shape) == (4, C_MODEL) and all(l.bias is not None for l in lins))
    return shapes_ok and all(p.dtype == torch.float32 for l in (class_embedding, *lins) for p in (l.weight, l.bias))


def _shadows(class_embedding: nn.Linear, lins, device) -> ShadowedLinears:
    sh = _SHADOWS.get(class_embedding)
    if sh is None:
        sh = ShadowedLinears()
        sh.register("cls", [class_embedding.weight], [class_embedding.bias], pad_rows=[_CLS_PAD])
        sh.register("m1", [lins[0].weight], [lins[0].bias])
        sh.register("m2", [lins[1].weight], [lins[1].bias])
        sh.register("m3", [lins[2].weight], [lins[2].bias], pad_rows=[_BOX_PAD])
        _SHADOWS[class_embedding] = sh
    sh.refresh(device)
    return sh


class _Heads(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n_cls, wc16, bc32, w1_16, b1_32, w2_16, b2_32, w3_16, b3_32, *params):
        lead, C = x.shape[:-1], x.shape[-1]
        x16 = G._mat(x.reshape(-1, C), "heads(x)")
        M, dev = x16.shape[0], x.device
        logits = torch.empty(M, n_cls, dtype=torch.float32, device=dev)
        boxes = torch.empty(M, 4, dtype=torch.float32, device=dev)
        y1 = torch.empty(M, C, dtype=torch.bfloat16, device=dev)
        y2 = torch.empty(M, C, dtype=torch.bfloat16, device=dev)
        G.gemm(x16, wc16[:n_cls], bias=bc32, out=logits)
        h1 = G.gemm(x16, w1_16, epilogue=G.EPI_GELU, bias=b1_32, aux=y1)
        h2 = G.gemm(h1, w2_16, epilogue=G.EPI_GELU, bias=b2_32, aux=y2)
        G.gemm(h2, w3_16[:4], epilogue=G.EPI_SIGMOID, bias=b3_32, out=boxes)
        ctx.save_for_backward(x16, y1, h1, y2, h2, boxes, wc16, w1_16, w2_16, w3_16)
        ctx.n_cls, ctx.x_dtype, ctx.lead = n_cls, x.dtype, tuple(lead)
        return logits.view(*lead, n_cls), boxes.view(*lead, 4)

    @staticmethod
    def backward(ctx, d_logits, d_boxes):
        x16, y1, h1, y2, h2, boxes, wc16, w1_16, w2_16, w3_16 = ctx.saved_tensors
        n_cls, M, dev = ctx.n_cls, x16.shape[0], x16.device
        dense = lambda g, n: (torch.zeros(M, n, dtype=torch.float32, device=dev) if g is None
                              else g.reshape(M, n).to(torch.float32).contiguous())
        d_logits, d_boxes = dense(d_logits, n_cls), dense(d_boxes, 4)
        dl16 = torch.empty(M, _CLS_PAD, dtype=torch.bfloat16, device=dev)
        dz16 = torch.empty(M, _BOX_PAD, dtype=torch.bfloat16, device=dev)
        _lib.call("detr_heads_grad_prep", d_logits.data_ptr(), n_cls, d_boxes.data_ptr(), boxes.data_ptr(), dl16.data_ptr(), _CLS_PAD,
                  dz16.data_ptr(), _BOX_PAD, M, _lib.stream_ptr())
        dwc, dbc = G.gemm_wgrad(dl16, x16)
        dw3, db3 = G.gemm_wgrad(dz16, h2)
        dh2 = G.gemm(dz16, w3_16, b_kn=True, epilogue=G.EPI_GELU_BWD, aux=y2)
        dw2, db2 = G.gemm_wgrad(dh2, h1)
        dh1 = G.gemm(dh2, w2_16, b_kn=True, epilogue=G.EPI_GELU_BWD, aux=y1)
        dw1, db1 = G.gemm_wgrad(dh1, x16)
        dx = None
        if ctx.needs_input_grad[0]:
            # the two branches meet in the residual epilogue: dx = dl16 . Wc + (dh1 . W1)
            dxb = G.gemm(dh1, w1_16, b_kn=True, out_dtype=ctx.x_dtype)
            dx = G.gemm(dl16, wc16, b_kn=True, epilogue=G.EPI_RES, res=dxb).view(*ctx.lead, x16.shape[1])
        return (dx, None, None, None, None, None, None, None, None, None,
                dwc[:n_cls], dbc[:n_cls], dw1, db1, dw2, db2, dw3[:4], db3[:4])


def predict(decoded: torch.Tensor, class_embedding: nn.Linear, bbox_embedding: nn.Module):
    """(pred_logits, pred_boxes) of detr/model.py:92-93 for decoder output `decoded` (..., 256)."""
    if not supported(decoded, class_embedding, bbox_embedding):
        return class_embedding(decoded), bbox_embedding(decoded).sigmoid()
    _lib.require_cuda(decoded, "heads.predict")
    lins = _mlp_linears(bbox_embedding)
    sh = _shadows(class_embedding, lins, decoded.device)
    (wc16, bc32), (w1, b1), (w2, b2), (w3, b3) = (sh.get_w_b32(k) for k in ("cls", "m1", "m2", "m3"))
    with torch.autocast("cuda", enabled=False):
        return _Heads.apply(decoded, class_embedding.out_features, wc16, bc32, w1, b1, w2, b2, w3, b3,
                            class_embedding.weight, class_embedding.bias, *[p for l in lins for p in (l.weight, l.bias)])
