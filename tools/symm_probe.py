#!/usr/bin/env python
"""Gradient all-reduce candidates at the model's real size (41.5 M fp32 = 166 MB): NCCL vs torch's symmetric-memory kernels
(NVLS multimem all-reduce through the NVSwitch, two-shot over peer pointers).  torchrun --nproc-per-node N tools/symm_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem  # noqa: E402

rank, world = dist.get_rank(), dist.get_world_size()
gname = dist.group.WORLD.group_name
N = 41_525_280


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / n, 4)


res = {"world": world}
plain = torch.randn(N, device=dev)
res["nccl_fp32_ms"] = t(lambda: dist.all_reduce(plain))
try:
    try:
        symm_mem.enable_symm_mem_for_group(gname)
    except Exception as e:  # newer versions enable it implicitly
        res["enable_note"] = repr(e)[:80]
    buf = symm_mem.empty(N, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(buf, gname)
    res["multicast"] = bool(getattr(hdl, "multicast_ptr", 0))
    # correctness on the real size: every rank contributes rank + 1
    for name, op in (("multimem_all_reduce_", lambda: torch.ops.symm_mem.multimem_all_reduce_(buf, "sum", gname)),
                     ("two_shot_all_reduce_", lambda: torch.ops.symm_mem.two_shot_all_reduce_(buf, "sum", gname))):
        try:
            buf.fill_(float(rank + 1))
            torch.cuda.synchronize(); dist.barrier()
            op()
            torch.cuda.synchronize(); dist.barrier()
            want = world * (world + 1) / 2
            ok = bool((buf[:1000] == want).all() and (buf[-1000:] == want).all() and buf[N // 2] == want)
            buf.normal_()
            res[name + "ok"] = ok
            res[name + "ms"] = t(op)
        except Exception as e:
            res[name + "error"] = repr(e)[:200]
except Exception as e:
    res["symm_mem_error"] = repr(e)[:300]
if rank == 0:
    print(res, flush=True)
dist.barrier()
dist.destroy_process_group()
