"""HungarianMatcher -- same constructor and forward signature as the reference (detr/matcher.py:17-99), executed by
ONE CUDA launch over all images (and, through `match_layers`, all decoder layers) with no host synchronisation.

Differences a maintainer should know:
  * indices are returned as int64 CUDA tensors (the reference returns CPU tensors after a `.cpu()` sync,
    detr/matcher.py:94-97); pass `return_cpu=True` to get the reference's placement.
  * data faults do not raise inside forward(): they set bits in `self.status` (device int32) and
    `check_status()` raises the reference's exception types at a sync point of the caller's choosing.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
from torch import nn

from . import _lib
from .targets import PackedTargets, pack_targets

ST_DEGENERATE_BOX, ST_INVALID_COST, ST_INFEASIBLE, ST_BAD_LABEL = 1, 2, 4, 8


def raise_for_status(bits: int) -> None:
    """Map device status bits to the exceptions the reference path raises."""
    if bits & ST_DEGENERATE_BOX:
        raise AssertionError("degenerate box: x2 < x1 or y2 < y1 (detr/utils.py:87-88)")
    if bits & ST_BAD_LABEL:
        raise IndexError("ground-truth label outside [0, num_classes]")
    if bits & ST_INVALID_COST:
        raise ValueError("matrix contains invalid numeric entries")
    if bits & ST_INFEASIBLE:
        raise ValueError("cost matrix is infeasible")


def _rows(t: torch.Tensor, row: int) -> torch.Tensor:
    """Make the last dim contiguous (and 16-byte aligned rows for boxes) without copying when already so."""
    if t.dtype != torch.float32:
        t = t.float()
    ok = t.stride(-1) == 1 and t.shape[-1] == row
    if row == 4:
        ok = ok and t.data_ptr() % 16 == 0 and all(s % 4 == 0 for s in t.stride()[:-1])
    return t if ok else t.contiguous()


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1, return_cpu: bool = False):
        super().__init__()
        self.cost_class = cost_class
        self.cost_bbox = cost_bbox
        self.cost_giou = cost_giou
        self.return_cpu = return_cpu
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs can't be 0"
        self._status = None

    # -- status word --------------------------------------------------------------------------------------
    def status_tensor(self, device: torch.device) -> torch.Tensor:
        if self._status is None or self._status.device != device:
            self._status = torch.zeros(1, dtype=torch.int32, device=device)
        return self._status

    def check_status(self) -> None:
        """Synchronising check; raises what the reference would have raised inside forward()."""
        if self._status is not None:
            bits = int(self._status.item())
            if bits:
                self._status.zero_()
                raise_for_status(bits)

    # -- kernels ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def match_layers(self, logits: torch.Tensor, boxes: torch.Tensor, pt: PackedTargets,
                     export_cost: bool = False):
        """logits (B,L,Q,K), boxes (B,L,Q,4) -> packed (idx_q, idx_gt) int64 [L*sum(n_b)] (+ packed costs).

        Problem (b,l) occupies [L*match_off[b] + l*n_b, +n_b) -- see include/detr_b200.h."""
        _lib.require_cuda(logits, "HungarianMatcher")
        B, L, Q, K = logits.shape
        logits, boxes = _rows(logits, K), _rows(boxes, 4)
        dev = logits.device
        n_out = L * sum(pt.n_match)
        idx = torch.empty(2, max(n_out, 1), dtype=torch.int64, device=dev)
        need_ws = _lib.load().detr_matcher_smem_bytes(Q, pt.max_count, 4) < 0
        cost = torch.empty(max(Q * L * pt.total, 1), dtype=torch.float32, device=dev) if (export_cost or need_ws) else None
        st = self.status_tensor(dev)
        order_ws = torch.empty(B, dtype=torch.int32, device=dev) if B * L > 4 * 148 else None   # largest problems first
        _lib.call(
            "detr_hungarian_match_f32",
            logits.data_ptr(), logits.stride(0), logits.stride(1), logits.stride(2),
            boxes.data_ptr(), boxes.stride(0), boxes.stride(1), boxes.stride(2),
            pt.labels.data_ptr(), pt.boxes.data_ptr(), pt.gt_off.data_ptr(), pt.match_off.data_ptr(),
            B, L, Q, K, pt.max_count, float(self.cost_class), float(self.cost_bbox), float(self.cost_giou),
            _lib.ptr(cost), idx[0].data_ptr(), idx[1].data_ptr(), st.data_ptr(), _lib.ptr(order_ws), _lib.stream_ptr())
        return (idx[0][:n_out], idx[1][:n_out], cost) if export_cost else (idx[0][:n_out], idx[1][:n_out])

    @torch.no_grad()
    def cost_matrices(self, logits: torch.Tensor, boxes: torch.Tensor, gt_labels: Sequence[torch.Tensor],
                      gt_boxes: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """Cost matrices only, one (Q, M_b) tensor per image (detr/matcher.py:66-93). logits (B,Q,K)."""
        _lib.require_cuda(logits, "HungarianMatcher")
        B, Q, K = logits.shape
        pt = pack_targets(gt_labels, gt_boxes, Q, logits.device)
        lg, bx = _rows(logits, K), _rows(boxes, 4)
        cost = torch.empty(max(Q * pt.total, 1), dtype=torch.float32, device=logits.device)
        st = self.status_tensor(logits.device)
        _lib.call(
            "detr_cost_matrix_f32",
            lg.data_ptr(), lg.stride(0), 0, lg.stride(1), bx.data_ptr(), bx.stride(0), 0, bx.stride(1),
            pt.labels.data_ptr(), pt.boxes.data_ptr(), pt.gt_off.data_ptr(), B, 1, Q, K, pt.max_count,
            float(self.cost_class), float(self.cost_bbox), float(self.cost_giou), cost.data_ptr(), st.data_ptr(),
            _lib.stream_ptr())
        out, o = [], 0
        for m in pt.counts:
            out.append(cost[o:o + Q * m].view(Q, m))
            o += Q * m
        return out

    @torch.no_grad()
    def forward(self, batch_pred_logits: torch.Tensor, batch_pred_boxes: torch.Tensor,
                batch_gt_labels: List[torch.Tensor], batch_gt_boxes: List[torch.Tensor]
                ) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        """Reference signature (detr/matcher.py:40-46): logits (B,Q,K), boxes (B,Q,4) cxcywh, per-image label and
        XYXY box lists -> list of (query_idx ascending, gt_idx), each of length min(Q, M_b)."""
        B, Q, _ = batch_pred_logits.shape
        pt = pack_targets(batch_gt_labels, batch_gt_boxes, Q, batch_pred_logits.device)
        iq, ig = self.match_layers(batch_pred_logits.unsqueeze(1), batch_pred_boxes.unsqueeze(1), pt)
        if self.return_cpu:
            iq, ig = iq.cpu(), ig.cpu()
            self.check_status()
        out, o = [], 0
        for n in pt.n_match:
            out.append((iq[o:o + n], ig[o:o + n]))
            o += n
        return out


def linear_sum_assignment_cuda(costs: Sequence[torch.Tensor], status: torch.Tensor | None = None):
    """Batched drop-in for scipy.optimize.linear_sum_assignment (detr/matcher.py:94) on CUDA cost matrices
    (float32 or float64, any shapes).  Returns a list of (row_ind, col_ind) int64 CUDA tensors."""
    if not costs:
        return []
    dev = costs[0].device
    _lib.require_cuda(costs[0], "linear_sum_assignment_cuda")
    dt = costs[0].dtype
    if dt not in (torch.float32, torch.float64) or any(c.dtype != dt or c.dim() != 2 for c in costs):
        raise ValueError("expected 2-D float32/float64 matrices of one dtype")
    flat = torch.cat([c.reshape(-1) for c in costs]) if sum(c.numel() for c in costs) else torch.zeros(1, dtype=dt, device=dev)
    nr = [int(c.shape[0]) for c in costs]
    nc = [int(c.shape[1]) for c in costs]
    n_out = [min(a, b) for a, b in zip(nr, nc)]
    meta32 = torch.tensor([nr, nc], dtype=torch.int32)
    sizes = torch.tensor([[a * b for a, b in zip(nr, nc)], n_out], dtype=torch.int64)
    meta64 = torch.zeros(2, len(costs), dtype=torch.int64)
    meta64[:, 1:] = sizes.cumsum(1)[:, :-1]
    meta32, meta64 = meta32.to(dev), meta64.to(dev)
    rows = torch.empty(max(sum(n_out), 1), dtype=torch.int64, device=dev)
    cols = torch.empty_like(rows)
    own_status = status is None
    if own_status:
        status = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.call("detr_lsap_f32" if dt == torch.float32 else "detr_lsap_f64", flat.data_ptr(), meta64[0].data_ptr(), meta32[0].data_ptr(), meta32[1].data_ptr(), len(costs),
            max(nr), max(nc), meta64[1].data_ptr(), rows.data_ptr(), cols.data_ptr(), status.data_ptr(),
            _lib.stream_ptr())
    if own_status:
        raise_for_status(int(status.item()))
    out, o = [], 0
    for n in n_out:
        out.append((rows[o:o + n], cols[o:o + n]))
        o += n
    return out
