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


class _PinnedRing:
    """`depth` pinned host staging tensors used round robin for async H2D copies.  A slot is rewritten only after the
    copy that last read it has executed (an event recorded behind that copy is synchronised first), so the host may
    run any number of steps ahead of the device without tearing a batch that is still waiting in the stream."""

    def __init__(self, shape, dtype, device: torch.device, depth: int = 3, fill=0):
        self.cuda = device.type == "cuda"
        self.slots = [torch.full(shape, fill, dtype=dtype) for _ in range(depth)]
        if self.cuda:
            self.slots = [t.pin_memory() for t in self.slots]
        self.events = [None] * depth
        self.i = -1

    def next(self) -> torch.Tensor:
        self.i = (self.i + 1) % len(self.slots)
        if self.events[self.i] is not None:
            self.events[self.i].synchronize()
        return self.slots[self.i]

    def submit(self, dst: torch.Tensor) -> None:
        """Copy the slot handed out by the last `next()` into `dst` on the current stream."""
        dst.copy_(self.slots[self.i], non_blocking=True)
        if self.cuda:
            if self.events[self.i] is None:
                self.events[self.i] = torch.cuda.Event()
            self.events[self.i].record(torch.cuda.current_stream())


class StaticTargets:
    """Packed targets in FIXED device buffers (capacity `cap` boxes per image) so that a CUDA graph captured once can
    be replayed on new ground truth: `update()` repacks on the host into pinned staging and issues async H2D copies;
    the kernels read the actual per-image counts from the device offsets.  The staging is a ring guarded by events
    (`_PinnedRing`): `update()` / `set_num_boxes()` never overwrite host memory an enqueued copy has yet to read."""

    def __init__(self, batch: int, num_queries: int, cap: int, device: torch.device, depth: int = 3):
        self.batch, self.num_queries, self.cap, self.device = batch, num_queries, cap, device
        n = batch * cap
        self._r_labels = _PinnedRing((n,), torch.int64, device, depth)
        self._r_boxes = _PinnedRing((n, 4), torch.float32, device, depth)
        self._r_offs = _PinnedRing((2, batch + 1), torch.int32, device, depth)
        self._r_nb = _PinnedRing((1,), torch.float32, device, depth, fill=1)
        self.labels = torch.zeros(n, dtype=torch.int64, device=device)
        self.boxes = torch.zeros(n, 4, device=device)
        self.offs = torch.zeros(2, batch + 1, dtype=torch.int32, device=device)
        self.num_boxes = torch.ones(1, device=device)      # normaliser read by the criterion kernels
        self.actual_counts = [0] * batch

    def update(self, gt_labels: Sequence[torch.Tensor], gt_boxes: Sequence[torch.Tensor]) -> int:
        """Repack (host tensors expected; device tensors are copied back first). Returns the local box count."""
        counts = [int(l.shape[0]) for l in gt_labels]
        if len(counts) != self.batch or max(counts, default=0) > self.cap:
            raise ValueError(f"StaticTargets(batch={self.batch}, cap={self.cap}) cannot hold counts {counts}")
        h_labels, h_boxes, h_offs = self._r_labels.next(), self._r_boxes.next(), self._r_offs.next()
        o = 0
        for l, b in zip(gt_labels, gt_boxes):
            m = l.shape[0]
            h_labels[o:o + m] = l.reshape(-1)
            h_boxes[o:o + m] = b.reshape(-1, 4)
            o += m
        c = torch.tensor(counts, dtype=torch.int32)
        h_offs[0, 1:] = c.cumsum(0)
        h_offs[1, 1:] = c.clamp(max=self.num_queries).cumsum(0)
        self._r_labels.submit(self.labels)
        self._r_boxes.submit(self.boxes)
        self._r_offs.submit(self.offs)
        self.actual_counts = counts
        return o

    def set_num_boxes(self, value: float) -> None:
        h = self._r_nb.next()
        h[0] = max(float(value), 1.0)
        self._r_nb.submit(self.num_boxes)

    def bytes_per_update(self) -> int:
        return sum(r.slots[0].numel() * r.slots[0].element_size() for r in (self._r_labels, self._r_boxes, self._r_offs, self._r_nb))

    @property
    def packed(self) -> PackedTargets:
        """Capacity-sized view for the launchers: host-side sizes are upper bounds, device offsets are exact."""
        return PackedTargets(self.labels, self.boxes, self.offs[0], self.offs[1], [self.cap] * self.batch,
                             [min(self.num_queries, self.cap)] * self.batch, self.num_queries)
