"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_targets(fx, device="cpu"):
    counts = fx["gt_counts"].tolist()
    labels = torch.from_numpy(fx["gt_labels"]).split(counts)
    boxes = torch.from_numpy(fx["gt_boxes"]).split(counts)
    return {"class_idx": [l.to(device) for l in labels], "boxes_normalized": [b.to(device) for b in boxes]}


def golden_losses(fx):
    return {k[len("loss/"):]: float(v) for k, v in fx.items() if k.startswith("loss/")}
