"""Packed ("CSR") ground truth: the reference passes per-image lists (detr/data.py:205-220 `class_idx`,
`boxes_normalized`; consumed at detr/loss.py:217 and detr/matcher.py:44-46); the kernels want them concatenated
with prefix offsets.  Lengths come from tensor shapes, so packing never synchronises with the device."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import torch


@dataclass
class PackedTargets:
    labels: torch.Tensor      # int64 [sumM]
    boxes: torch.Tensor       # float32 [sumM, 4] XYXY normalised
    gt_off: torch.Tensor      # int32 [B+1] device
    match_off: torch.Tensor   # int32 [B+1] device, prefix of min(Q, M_b)
    counts: list              # host ints M_b
    n_match: list             # host ints min(Q, M_b)
    num_queries: int

    @property
    def batch(self) -> int:
        return len(self.counts)

    @property
    def total(self) -> int:
        return sum(self.counts)

    @property
    def max_count(self) -> int:
        return max(self.counts) if self.counts else 0


def pack_targets(gt_labels: Sequence[torch.Tensor], gt_boxes: Sequence[torch.Tensor], num_queries: int,
                 device: torch.device) -> PackedTargets:
    if len(gt_labels) != len(gt_boxes):
        raise ValueError("gt_labels and gt_boxes must have one entry per image")
    counts = [int(l.shape[0]) for l in gt_labels]
    for c, bx in zip(counts, gt_boxes):
        if bx.shape[0] != c or (c and bx.shape[-1] != 4):
            raise ValueError("gt_boxes[i] must be (len(gt_labels[i]), 4)")
    n_match = [min(num_queries, c) for c in counts]
    if sum(counts):
        labels = torch.cat([l.reshape(-1) for l in gt_labels]).to(device=device, dtype=torch.int64)
        boxes = torch.cat([b.reshape(-1, 4) for b in gt_boxes]).to(device=device, dtype=torch.float32).contiguous()
    else:
        labels = torch.zeros(1, dtype=torch.int64, device=device)[:0]
        boxes = torch.zeros(1, 4, dtype=torch.float32, device=device)[:0]
    offs = torch.zeros(2, len(counts) + 1, dtype=torch.int32)
    offs[0, 1:] = torch.tensor(counts, dtype=torch.int32).cumsum(0)
    offs[1, 1:] = torch.tensor(n_match, dtype=torch.int32).cumsum(0)
    if device.type == "cuda":
        offs = offs.pin_memory().to(device, non_blocking=True)
    return PackedTargets(labels, boxes, offs[0], offs[1], counts, n_match, num_queries)
