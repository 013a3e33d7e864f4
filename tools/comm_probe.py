#!/usr/bin/env python
"""Times the pieces of the data-parallel gradient exchange of bench.py (flatten, NCCL all-reduce, un-flatten + average) at the
model's real gradient sizes.  torchrun --nproc-per-node N tools/comm_probe.py"""
import os, sys, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "detr-object-detection_b200"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from detr_b200.harness import DetrHarness
from detr_b200.model import DETRConfig
model = DetrHarness(DETRConfig(num_classes=91)).to(dev)
params = [p for p in model.parameters() if p.requires_grad]
grads = [torch.randn_like(p) for p in params]
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
flat = torch.cat([g.reshape(-1) for g in grads])
res = {"numel_M": flat.numel() / 1e6,
       "cat_ms": t(lambda: torch.cat([g.reshape(-1) for g in grads])),
       "allreduce_fp32_ms": t(lambda: dist.all_reduce(flat)),
       "allreduce_bf16_ms": t(lambda: dist.all_reduce(flat.bfloat16())),
       "unflatten_div_ms": t(lambda: (torch._foreach_copy_(grads, [v.view_as(g) for v, g in zip(torch.split(flat, [g.numel() for g in grads]), grads)]), torch._foreach_div_(grads, 2.0)))}
for mb in (8, 32, 64):
    x = torch.empty(mb * 1024 * 1024 // 4, device=dev)
    res[f"allreduce_{mb}MB_ms"] = t(lambda: dist.all_reduce(x))
if dist.get_rank() == 0:
    print(res, flush=True)
dist.destroy_process_group()
