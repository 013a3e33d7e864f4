#!/usr/bin/env python
"""Debugging aid: one graphed training step with the weight-gradient side stream on and off; per-parameter gradient differences."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "detr-object-detection_b200"))
from detr_b200 import HungarianMatcher, SetCriterion  # noqa: E402
from detr_b200.harness import DetrHarness, GraphedTrainStep, make_optimizer, synthetic_batch  # noqa: E402
from detr_b200.model import DETRConfig  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = DETRConfig(num_classes=11, num_object_queries=20, num_encoder_layers=2, num_decoder_layers=2)
m0 = DetrHarness(cfg).to(dev).to(memory_format=torch.channels_last)
m0.train(False)
c0 = SetCriterion(11, HungarianMatcher(1.0, 5.0, 2.0)).to(dev)
b = synthetic_batch(2, 160, 200, 11, 6, seed=1)
res = {}
for tag, side, heads in (("off", "0", ""), ("on", "1", ""), ("on2", "1", "")):
    os.environ["DETR_B200_WGRAD_STREAM"] = side
    os.environ["DETR_DBG"] = heads
    m, c = copy.deepcopy(m0), copy.deepcopy(c0)
    o = make_optimizer(m, lr=1e-4, capturable=True)
    g = GraphedTrainStep(m, c, o, b, gt_cap=8, warmup=2)
    g.load(b)
    loss = float(g.step())
    torch.cuda.synchronize()
    names = {id(p): n for n, p in m.named_parameters()}
    res[tag] = (loss, {names[id(p)]: v.detach().clone() for p, v in zip(g.fopt.params, g.fopt.grad_views)})
    print(tag, "loss", loss)
ref = res["off"][1]
for tag in ("on", "on2"):
    bad = []
    for n, v in res[tag][1].items():
        d = (v - ref[n]).abs().max().item()
        s = ref[n].abs().max().item()
        if d > 1e-3 * s + 1e-9:
            bad.append((d / (s + 1e-12), n, d, s))
    bad.sort(reverse=True)
    print(tag, "parameters that differ from 'off':", len(bad))
    for r, n, d, s in bad[:25]:
        print(f"   {n:70s} diff {d:.3e} scale {s:.3e}")
