#!/usr/bin/env python
"""torch.profiler view of ONE eager training step of bench.py's configuration: ATen operators that launch copy / fill / add
kernels, grouped by (operator, input shapes, Python call site).  Finds glue launches around the library's kernels.
usage: python tools/step_op_profile.py [pattern ...]   (default patterns: copy_ fill_ add zeros contiguous)"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "detr-object-detection_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from detr_b200 import HungarianMatcher, SetCriterion
from detr_b200.harness import DetrHarness, batch_to, make_optimizer, synthetic_batch, train_step
from detr_b200.model import DETRConfig

pats = sys.argv[1:] or ["aten::copy_", "aten::fill_", "aten::add", "aten::zero_", "aten::threshold_backward", "aten::sum", "aten::cat", "aten::stack"]
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = DetrHarness(DETRConfig(num_classes=91)).to(dev).to(memory_format=torch.channels_last).train()
crit = SetCriterion(91, HungarianMatcher(1.0, 5.0, 2.0), 1.0, 5.0, 2.0, 0.1).to(dev).train()
opt = make_optimizer(model)
host = synthetic_batch(8, 800, 1066, 91, 20, seed=100, pin=True)
host["image"] = host["image"].contiguous(memory_format=torch.channels_last)
batch = batch_to(host, dev)
for _ in range(3):
    train_step(model, crit, opt, batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    train_step(model, crit, opt, batch)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if not any(p in e.name for p in pats):
        continue
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0.0)
    if t <= 0:
        continue
    site = ""
    for fr in (e.stack or []):
        if "detr_b200" in fr or "bench.py" in fr:
            site = fr.split("detr-object-detection_b200/")[-1][:70]
            break
    agg[(e.name, str(e.input_shapes)[:90], site)][0] += 1
    agg[(e.name, str(e.input_shapes)[:90], site)][1] += t
tot = sum(v[1] for v in agg.values())
print(f"matched operators: {tot:.0f} us of device time in one step (includes children)")
for (n, sh, site), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{t:8.1f} us {c:4d}  {n:28s} {sh:90s} {site}")
