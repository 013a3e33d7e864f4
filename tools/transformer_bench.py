#!/usr/bin/env python
"""Times Encoder+Decoder(+heads+criterion) fwd+bwd alone at BASELINE config 2 shapes (B=8, S=850, Q=100) under bf16
autocast, train mode, and splits the time into the library's own launches vs everything else (ATen/cuBLAS)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "detr-object-detection_b200"))
import torch
from detr_b200 import _lib
from detr_b200.model import DETRConfig, Decoder, Encoder

dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = DETRConfig(num_classes=91)
enc, dec = Encoder(cfg).to(dev).train(), Decoder(cfg).to(dev).train()
B, S, Q = 8, 850, 100
x = torch.randn(B, S, 256, device=dev, dtype=torch.bfloat16, requires_grad=True)
pos = torch.randn(B, S, 256, device=dev)
qe = torch.randn(B, Q, 256, device=dev)
mask = torch.zeros(B, S, dtype=torch.bool, device=dev)

def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        mem = enc(x, pos, mask)
        out = dec(mem, pos, qe, mask)
    out.float().sum().backward()

for _ in range(5): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): step()
b.record(); torch.cuda.synchronize()
tot = a.elapsed_time(b) / 10
with _lib.profile() as prof:
    for _ in range(5): step()
torch.cuda.synchronize()
own = sum(t for _, t in prof.summary().values()) / 5
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as p:
    step(); torch.cuda.synchronize()
rows = sorted(p.key_averages(), key=lambda e: -e.device_time_total)[:25]
print(json.dumps({"transformer_fwd_bwd_ms": round(tot, 3), "own_kernels_ms": round(own, 3), "other_ms": round(tot - own, 3)}))
for e in rows:
    print(f"{e.device_time_total/1e3:8.3f} ms x{e.count:4d}  {e.key[:100]}")
