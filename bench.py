#!/usr/bin/env python
"""bench.py -- DETR-R50 training-step throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

b200 arm (default): BASELINE config 2 -- full training step (forward, HungarianMatcher + SetCriterion, backward,
gradient all-reduce, clip 1.0, AdamW) in bf16 autocast, 8 synthetic 800x1066 images per GPU, random-init DETR-R50,
train mode (dropout on).  The transformer attention, the matcher and the criterion run on libdetr_b200.so; the
ResNet-50 backbone, the GEMM projections and the heads are cuDNN/cuBLAS through PyTorch (out of scope).
One JSON line is printed by rank 0: `value` with inputs resident in HBM, `e2e` through the public step with pinned-host
inputs copied in and the loss read back every step, `roofline` of the dominant own kernel measured live with CUDA
events, and `cpu_baseline` (the oracle port of the reference on the host cores, bounded sample).

reference arm: the reference's own CPU implementation of the path -- here the oracle port (oracle/detr_oracle.py,
pinned to the real reference by tests/golden) -- timed on the host cores; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "detr-object-detection_b200"))

import torch  # noqa: E402

METRIC = "detr_r50_train_images_per_sec"
UNIT = "images/s"
PER_GPU_BATCH = 8
NUM_CLASSES = 91
H, W = 800, 1066
SP = (H // 32) * ((W + 31) // 32)  # 850 tokens
WORKLOAD = "DETR-R50 full training step bf16, 8 img/GPU synthetic 800x1066 (padded 800x1088, 850 tokens), 100 queries, 91 classes"
# BASELINE.json configs that are bench lines: 2 (the headline metric, what the driver runs), 5 (R101, 300 queries, 4 img/GPU) and
# 4 (DC5: the 6+6 layer transformer on synthetic stride-16 tokens, 3 350 per image -- the reference's Backbone hard-codes stride 32,
# detr/model.py:431-435, so DC5 exists only as the transformer workload of SURVEY.md 8d)
CONFIGS = {
    2: {"backbone": "resnet50", "queries": 100, "per_gpu_batch": 8, "metric": METRIC, "workload": WORKLOAD},
    5: {"backbone": "resnet101", "queries": 300, "per_gpu_batch": 4, "metric": "detr_r101_q300_train_images_per_sec",
        "workload": "DETR-R101 full training step bf16, 300 object queries, 4 img/GPU synthetic 800x1066 (BASELINE config 5)"},
    4: {"backbone": None, "queries": 100, "per_gpu_batch": 8, "metric": "detr_dc5_transformer_fwd_bwd_images_per_sec",
        "workload": "DETR-R50-DC5 transformer (6 encoder + 6 decoder layers, 100 queries) forward + backward on synthetic stride-16 "
                    "tokens: 50 x 67 = 3 350 per image, 8 img/GPU, bf16 autocast, train mode (BASELINE config 4)"},
}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"tflops_burst": p["bf16_tflops"], "tflops_sustained": p["bf16_tflops_sustained"], "hbm": p["hbm_gbs"], "src": "measured"}
    except Exception:
        return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).  NVML is polled
    from a thread every 5 ms (an `nvidia-smi -lms` child needs longer to start than a short timed region lasts);
    `nvidia-smi --query-gpu` is the fallback when the NVML binding cannot be loaded."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.sm, self.mx, self.reasons, self.src = index, [], None, set(), None
        self._stop = threading.Event()
        self._thread = None

    def _nvml_loop(self):
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.index
        if vis:
            try:
                idx = int(vis.split(",")[self.index])
            except (ValueError, IndexError):
                pass
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        self.src = "nvml"
        while not self._stop.is_set():
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for name, bit in bits.items():
                if r & bit:
                    self.reasons.add(name)
            self._stop.wait(0.005)

    def _smi_loop(self):
        self.src = "nvidia-smi"
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=10).stdout
            except Exception:
                return
            f = [x.strip() for x in out.strip().split(",")]
            if len(f) >= 6:
                try:
                    self.sm.append(float(f[0])); self.mx = float(f[1])
                except ValueError:
                    pass
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)

    def _run(self):
        try:
            self._nvml_loop()
        except Exception:
            self._smi_loop()

    def start(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=15)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": ["no clock samples"], "samples": 0, "source": self.src}
        return {"sm_mhz": statistics.median(self.sm), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": self.src}


# ------------------------------------------------------------------------------------------------ b200 arm
def run_b200(args):
    import torch.distributed as dist
    from detr_b200 import HungarianMatcher, SetCriterion, _lib
    from detr_b200.harness import DetrHarness, GraphedTrainStep, batch_bytes, batch_to, make_optimizer, synthetic_batch, train_step
    from detr_b200.model import DETRConfig

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if args.gpus != 1:
            raise SystemExit(f"--gpus {args.gpus} needs a torchrun launch with {args.gpus} ranks (WORLD_SIZE={world})")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (b200 arm) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if not args.eager and os.environ.get("DETR_B200_OVERLAP_AR", "0") == "1":
            # the bucketed all-reduces are CAPTURED in the forward/backward graph: as for DistributedDataParallel under CUDA graphs
            # (PyTorch notes, "Usage with DistributedDataParallel"), NCCL's asynchronous error handling must be off -- its watchdog
            # thread polls CUDA events, which is not allowed while a capture is in progress
            os.environ["TORCH_NCCL_ASYNC_ERROR_HANDLING"] = "0"
            os.environ["NCCL_ASYNC_ERROR_HANDLING"] = "0"
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1234 + rank)

    C = CONFIGS[args.config]
    if args.config == 4:
        return run_dc5(args, dev, world, rank, C)
    global PER_GPU_BATCH
    PER_GPU_BATCH = C["per_gpu_batch"]
    cfg = DETRConfig(num_classes=NUM_CLASSES, backbone=C["backbone"], num_object_queries=C["queries"])
    torch.manual_seed(1234)   # identical replicas on every rank (data differs per rank)
    model = DetrHarness(cfg).to(dev).to(memory_format=torch.channels_last).train()
    crit = SetCriterion(NUM_CLASSES, HungarianMatcher(cost_class=1.0, cost_bbox=5.0, cost_giou=2.0), 1.0, 5.0, 2.0, 0.1).to(dev).train()
    torch.manual_seed(1234 + rank)

    host = synthetic_batch(PER_GPU_BATCH, H, W, NUM_CLASSES, 20, seed=100 + rank, pin=True)
    host["image"] = host["image"].contiguous(memory_format=torch.channels_last).pin_memory()

    if args.eager:
        # eager path: DDP bucketed all-reduce overlapped with backward, one Python-launched kernel at a time
        step_model = model
        if world > 1:
            step_model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)
        opt = make_optimizer(step_model)
        resident = batch_to(host, dev)
        torch.cuda.synchronize()

        def step_resident():
            return train_step(step_model, crit, opt, resident)

        def step_e2e():
            b = batch_to(host, dev, non_blocking=True)     # H2D of this step's inputs from pinned memory
            return float(train_step(step_model, crit, opt, b).item())   # D2H read of the step's loss
        h2d = batch_bytes(host)
        own_launches = None
    else:
        # default: the same step captured in CUDA graphs (harness.GraphedTrainStep) -- replay is GPU-bound
        opt = make_optimizer(model, capturable=True)
        graphed = GraphedTrainStep(model, crit, opt, host, accumulate=args.accum,
                                   # opt-in: measured at 2 GPUs 12.83 ms (bucketed, overlapped, captured) vs 12.74 ms (one all-reduce
                                   # between the graphs): the NCCL kernels take SMs from the backward they overlap with
                                   overlap_allreduce=os.environ.get("DETR_B200_OVERLAP_AR", "0") == "1")
        h2d = graphed.load(host)
        torch.cuda.synchronize()

        def step_resident():
            for _ in range(args.accum - 1):       # non-boundary micro-steps: forward/backward only (no collective, no update)
                graphed.step()
            return graphed.step()

        graphed.prefetch(host)

        def step_e2e():
            # double-buffered input pipeline, as a data loader feeds a training loop: this step consumes the batch whose H2D
            # copy (pinned host memory -> staging buffer, copy stream) was started during the previous step, and starts the
            # copy for the next one -- one full-batch H2D copy and one loss read-back inside every timed step
            for _ in range(args.accum):
                graphed.commit()                            # wait for the staged images, D2D into the graph's input, GT upload
                graphed.prefetch(host)                      # H2D of the next micro-batch's images, overlapping this one's compute
                loss = graphed.step()
            return float(loss.item())                       # D2H read of the step's loss
        own_launches = graphed.own_launches_per_step * args.accum
        h2d *= args.accum

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    if args.ncu_step:
        # profiling aid (never a bench number): `ncu --profile-from-start off ... bench.py --ncu-step` captures one step
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count
    ms = timed(step_resident, args.steps)
    launches = (_lib.launch_count - l0) if own_launches is None else own_launches * args.steps
    clocks = sampler.stop() if rank == 0 else None
    crit.check_status()
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # ---- live per-kernel timing of the library's launches (CUDA events on the launching stream), 3 extra steps ----
    # (graph replays cannot be instrumented from the host, so these 3 steps run the SAME kernels eagerly)
    if args.eager:
        prof_step = step_resident
    else:
        eager_opt = make_optimizer(model)
        resident = batch_to(host, dev)
        prof_step = lambda: train_step(model, crit, eager_opt, resident)
        # (at N > 1 these three untimed steps skip the gradient all-reduce: the replicas diverge AFTER every timed region
        #  has ended, which is harmless; every rank runs them so that the num_boxes all-reduce stays collective)
    prof_step()                                  # (the eager step's own warm-up: allocator, autograd graph)
    torch.cuda.synchronize()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with _lib.profile() as prof:
        pe0.record()
        for _ in range(3):
            prof_step()
        pe1.record()
    torch.cuda.synchronize()
    rows = prof.summary()
    eager_step_ms = pe0.elapsed_time(pe1) / 3
    step_ms = ms / args.steps

    if rank == 0:
        imgs = PER_GPU_BATCH * world * args.accum
        pk = peaks()
        own = {f"{n}{list(t) if t else ''}": {"calls_per_step": c / 3, "ms_per_step": round(t_ms / 3, 4)} for (n, t), (c, t_ms) in sorted(rows.items(), key=lambda kv: -kv[1][1])}
        # dominant own launch: encoder self-attention backward (dK/dV + dQ + delta), shape (B=8, nh=8, L=S=850)
        key_b = ("detr_attention_bwd_bf16", (PER_GPU_BATCH, 8, SP, SP))
        key_f = ("detr_attention_fwd_bf16", (PER_GPU_BATCH, 8, SP, SP))
        roof = None
        if key_b in rows:
            calls, tot = rows[key_b]
            # algorithmic FLOPs per call: backward = 2.5 x forward core, forward core = 4*L*S*C per image
            flops = 2.5 * 4.0 * SP * SP * 256 * PER_GPU_BATCH
            ach = flops / (tot / calls * 1e-3) / 1e12
            roof = {"kernel": "attention_bwd (delta + fused dK/dV/dQ-partial + dQ reduction), encoder self-attention B=8 nh=8 L=S=850", "bound": "tensor",
                    "achieved": round(ach, 2), "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": round(ach / pk["tflops_sustained"], 4),
                    "peak_source": pk["src"] + " sustained (kernel timed inside a long step)",
                    # dram__bytes_read.sum + dram__bytes_write.sum of the call's kernels from the committed `ncu --set full` capture
                    # (written by tools/ncu_traffic.py; ncu flushes L2 between kernels, so what the dQ reduction re-reads from L2 in a
                    # real step counts as DRAM traffic there); null when no capture of this build is committed
                    "traffic": ncu_traffic("attention_bwd_encoder"), "traffic_algorithmic": 14.1e6,
                    "ms_per_launch": round(tot / calls, 4), "flops_per_launch": flops,
                    "note": "head_dim 32: one MUFU exp per 128 tensor FLOPs caps the tensor pipe at ~25% (SURVEY.md 7.1); the persistent "
                            "kernel is bound by per-warp latency chains and issue slots, not by a pipe (profiles/r01_attention_ncu_full_v4.md)"}
            if key_f in rows:
                c2, t2 = rows[key_f]
                f2 = 4.0 * SP * SP * 256 * PER_GPU_BATCH
                roof["forward"] = {"ms_per_launch": round(t2 / c2, 4), "achieved": round(f2 / (t2 / c2 * 1e-3) / 1e12, 2),
                                   "frac": round(f2 / (t2 / c2 * 1e-3) / 1e12 / pk["tflops_sustained"], 4)}
        own_ms = sum(t for _, t in rows.values()) / 3
        out = {
            "metric": C["metric"], "value": round(imgs * args.steps / (ms * 1e-3), 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(step_ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "impl": "b200",
            "config": {"workload": C["workload"], "baseline_config": args.config, "global_batch": imgs, "per_gpu_batch": PER_GPU_BATCH,
                       "gradient_accumulation": args.accum, "parallelism": f"dp{world}",
                       "mode": "train (dropout on)", "optimizer": "AdamW fused, clip 1.0",
                       "execution": "eager + DDP" if args.eager else
                                    ("CUDA graphs (fwd+bwd with bucketed NCCL all-reduces overlapped on a side stream | clip+AdamW)"
                                     if (world > 1 and getattr(graphed, "_buckets", None) is not None) else
                                     "CUDA graphs (fwd+bwd | NCCL all-reduce of flat grads | clip+AdamW)"),
                       "l2": "no explicit flush: every step streams >1 GB of ResNet activations through the 126 MB L2",
                       "e2e_input_pipeline": "eager: blocking H2D per step" if args.eager else
                                             "double-buffered: the H2D copy of batch i+1 (copy stream) overlaps the compute of step i"},
            "e2e": {"value": round(imgs * args.steps / (ms_e2e * 1e-3), 3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof,
            # event-timed launches of libdetr_b200.so in the EAGER replay of the step, as a share of that same eager step (host-launch
            # bound, ~3x the graph replay); the share of the graph-replayed step is the ncu launch-list share in profiles/
            "own_kernels": {"ms_per_step": round(own_ms, 3), "eager_step_ms": round(eager_step_ms, 3),
                            "share_of_eager_step": round(own_ms / eager_step_ms, 4), "ncu_share_of_step": ncu_traffic("own_share_of_step"),
                            "by_call": own},
        }
        if world == 1 and args.config == 2:
            m = matcher_metric(dev)
            out["roofline_matcher"] = m.pop("roofline_matcher")
            out["roofline_criterion"] = m.pop("roofline_criterion")
            out["matcher"] = m
            out["cpu_baseline"] = cpu_baseline(sample_steps=8)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_dc5(args, dev, world, rank, C):
    """BASELINE config 4: the 6 + 6 layer transformer on stride-16 tokens (50 x 67 = 3 350 per image), forward + backward captured
    in one CUDA graph.  No backbone (the reference's is stride 32 only), no optimizer: this is the attention-dominated workload."""
    import torch.distributed as dist
    from detr_b200 import _lib, attention, gemm as gemm_mod
    from detr_b200.harness import positional_encoding_tokens
    from detr_b200.model import DETRConfig, Decoder, Encoder
    eh, ew = (int(v) for v in args.grid.split("x"))       # default 50 x 67 (stride 16); 25x34 = the stride-32 map of config 2
    B, Q = C["per_gpu_batch"], C["queries"]
    S = eh * ew
    torch.manual_seed(1234)
    cfg = DETRConfig(num_classes=NUM_CLASSES, num_object_queries=Q)
    enc, dec = Encoder(cfg).to(dev).train(), Decoder(cfg).to(dev).train()
    qe = torch.nn.Parameter(0.02 * torch.randn(Q, 256, device=dev))
    params = list(enc.parameters()) + list(dec.parameters()) + [qe]
    heights = torch.full((B,), 800, dtype=torch.int32, device=dev)
    widths = torch.full((B,), 1066, dtype=torch.int32, device=dev)
    pos, mask = positional_encoding_tokens(eh, ew, heights, widths, 16 if eh >= 50 else 32, 128, 10000)
    host_x = torch.randn(B, S, 256).bfloat16().pin_memory()        # what input_proj hands the encoder under autocast
    x = torch.empty(B, S, 256, dtype=torch.bfloat16, device=dev)
    x.copy_(host_x)
    w = torch.randn(B, 6, Q, 256, device=dev)
    loss_buf = torch.zeros((), device=dev)
    step_counter = torch.zeros(1, dtype=torch.int64, device=dev)
    attention.set_dropout_step_tensor(step_counter)

    def fwd_bwd():
        for q in params:
            q.grad = None
        step_counter.add_(1)
        xin = x.detach().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            mem = enc(xin, position_embedding=pos, key_padding_mask=mask)
            out = dec(mem, position_embedding=pos, object_query_embedding=qe[None].expand(B, -1, -1), key_padding_mask=mask)
        loss = (out.float() * w).mean()
        # weight-gradient GEMMs on a parallel branch of the graph (detr_b200/gemm.py), as in GraphedTrainStep
        prev = gemm_mod.wgrad_side_stream(os.environ.get("DETR_B200_WGRAD_STREAM", "1") != "0")
        try:
            loss.backward()
        finally:
            gemm_mod.wgrad_side_stream(prev)
        loss_buf.copy_(loss.detach())

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fwd_bwd()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    l0 = _lib.launch_count
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fwd_bwd()
    own_launches = _lib.launch_count - l0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_e2e():
        x.copy_(host_x, non_blocking=True)
        graph.replay()
        return float(loss_buf.item())

    for _ in range(max(args.warmup, 3)):
        graph.replay()
    if args.ncu_step:   # profiling aid (never a bench number): one replay between cudaProfilerStart/Stop
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        graph.replay()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    if args.graph_profile:
        # profiling aid: a second capture with timable event-record nodes around every C-ABI call: per-call time INSIDE the replay
        with _lib.profile(external=True) as gp:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                fwd_bwd()
        for _ in range(3):
            g2.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g2.replay(); e1.record()
        torch.cuda.synchronize()
        agg = {}
        for name, tag, a, b in gp.records:
            k = f"{name}{list(tag) if tag else ''}"
            n, t = agg.get(k, (0, 0.0))
            agg[k] = (n + 1, t + a.elapsed_time(b))
        tot = sum(t for _, t in agg.values())
        print(f"graph replay {e0.elapsed_time(e1):.3f} ms; C-ABI calls inside it: {tot:.3f} ms in {sum(n for n, _ in agg.values())} calls")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"  {t * 1e3:8.1f} us {n:4d} calls {t * 1e3 / n:7.1f} us/call  {k}")
        return
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    ms = timed(graph.replay, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, args.steps)
    with _lib.profile() as prof:
        for _ in range(3):
            fwd_bwd()
    torch.cuda.synchronize()
    rows = prof.summary()
    if rank == 0:
        pk = peaks()
        imgs = B * world
        roof = None
        kb, kf = ("detr_attention_bwd_bf16", (B, 8, S, S)), ("detr_attention_fwd_bf16", (B, 8, S, S))
        if kb in rows:
            calls, tot = rows[kb]
            flops = 2.5 * 4.0 * S * S * 256 * B
            ach = flops / (tot / calls * 1e-3) / 1e12
            roof = {"kernel": f"attention_bwd, DC5 encoder self-attention B={B} nh=8 L=S={S}", "bound": "tensor", "achieved": round(ach, 2),
                    "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": round(ach / pk["tflops_sustained"], 4),
                    "peak_source": pk["src"] + " sustained", "ms_per_launch": round(tot / calls, 4), "flops_per_launch": flops,
                    "traffic": ncu_traffic("attention_bwd_dc5")}
            if kf in rows:
                c2, t2 = rows[kf]
                f2 = 4.0 * S * S * 256 * B
                roof["forward"] = {"ms_per_launch": round(t2 / c2, 4), "achieved": round(f2 / (t2 / c2 * 1e-3) / 1e12, 2),
                                   "frac": round(f2 / (t2 / c2 * 1e-3) / 1e12 / pk["tflops_sustained"], 4)}
        own = {f"{n}{list(t) if t else ''}": {"calls_per_step": c / 3, "ms_per_step": round(t_ms / 3, 4)}
               for (n, t), (c, t_ms) in sorted(rows.items(), key=lambda kv: -kv[1][1])}
        print(json.dumps({
            "metric": C["metric"], "value": round(imgs * args.steps / (ms * 1e-3), 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "impl": "b200",
            "config": {"workload": C["workload"], "baseline_config": 4, "global_batch": imgs, "per_gpu_batch": B, "parallelism": f"replicas x{world}",
                       "tokens_per_image": S, "execution": "one CUDA graph (forward + backward)",
                       "l2": "no explicit flush: one step streams ~2 GB of activations through the 126 MB L2"},
            "e2e": {"value": round(imgs * args.steps / (ms_e2e * 1e-3), 3), "unit": UNIT, "h2d_bytes_per_step": host_x.numel() * 2,
                    "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": own_launches * args.steps, "clocks": clocks, "roofline": roof, "own_kernels": {"by_call": own},
        }), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ncu_traffic(key):
    """Numbers read from the committed ncu captures (profiles/r02_ncu_numbers.json, written by tools/ncu_traffic.py): never a
    literal in this file.  None when the key is absent."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_numbers.json"))).get(key)
    except Exception:
        return None


def matcher_metric(dev):
    """BASELINE.json's second metric, "matcher images/sec": config 3 (batch 256, 6 decoder layers, 100 queries, 92 logits,
    1..100 GT boxes per image, fp32) through HungarianMatcher (one fused cost + assignment launch for the 1 536 problems)
    and through SetCriterion forward + backward (matcher + criterion kernels).  CUDA events, 3 warm-ups, median of 10."""
    from detr_b200 import HungarianMatcher, SetCriterion, pack_targets
    B, L, Q, NC = 256, 6, 100, 91
    g = torch.Generator().manual_seed(1)
    logits = torch.randn(B, L, Q, NC + 1, generator=g).to(dev)
    boxes = torch.randn(B, L, Q, 4, generator=g).sigmoid().to(dev)
    counts = torch.randint(1, 101, (B,), generator=g).tolist()
    labels, gts = [], []
    for m in counts:
        c = torch.rand(m, 2, generator=g) * 0.6 + 0.2
        sz = torch.rand(m, 2, generator=g) * 0.30 + 0.02
        gts.append(torch.cat([c - sz / 2, c + sz / 2], dim=1).to(dev))
        labels.append(torch.randint(0, NC, (m,), generator=g, dtype=torch.int64).to(dev))
    matcher = HungarianMatcher(cost_class=1.0, cost_bbox=5.0, cost_giou=2.0)
    pt = pack_targets(labels, gts, Q, dev)

    def med(fn, n=10):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    ms_match = med(lambda: matcher.match_layers(logits, boxes, pt))
    crit = SetCriterion(NC, matcher, 1.0, 5.0, 2.0, 0.1).to(dev)
    lg, bx = logits.clone().requires_grad_(True), boxes.clone().requires_grad_(True)
    tg = {"class_idx": labels, "boxes_normalized": gts}

    def full():
        lg.grad = None; bx.grad = None
        out = crit({"pred_logits": lg, "pred_boxes": bx}, tg)
        sum(v for k, v in out.items() if k.startswith("loss")).backward()
    ms_full = med(full)
    # per-kernel CUDA-event timings of the criterion launches in the same loop (the C-ABI calls are bracketed on their stream)
    from detr_b200 import _lib
    for _ in range(2):
        full()
    with _lib.profile() as prof:
        for _ in range(10):
            full()
    torch.cuda.synchronize()
    kms = {n: v for (n, _), v in prof.median().items()}
    pk = peaks()
    by_m = sum(38400 + 40 * c for c in counts) * L                    # logits + boxes + GT in, indices out; costs stay in smem
    by_f = sum(36800 + 40 * c for c in counts) * L                    # criterion forward: logits + matched boxes / labels
    by_b = sum(2 * 36800 + 1600 + 40 * c for c in counts) * L         # backward: logits re-read, d_logits + d_boxes written
    def roof(name, byt, ms_k, note):
        ach = byt / (ms_k * 1e-3) / 1e9
        return {"kernel": name, "bound": "hbm", "achieved": round(ach, 1), "peak": pk["hbm"], "unit": "GB/s", "frac": round(ach / pk["hbm"], 4),
                "peak_source": pk["src"] + " burst (kernel timed alone)", "ms_per_launch": round(ms_k, 4), "bytes_per_launch": byt,
                "traffic": ncu_traffic(name.split()[0]), "note": note}
    roof_m = roof("hungarian_match_kernel (cost matrix + assignment, 1 536 problems)", by_m, ms_match,
                  "latency-bound by construction: serial augmenting paths (~500 dependent row scans per problem); reported against HBM for transparency")
    roof_c = {"forward": roof("criterion_fwd (dense kernel + finalize)", by_f, kms.get("detr_criterion_fwd_f32", float("nan")),
                              "two launches; ~17 us latency floor (offset -> index -> label/box -> class weight load chain per CTA)"),
              "backward": roof("criterion_bwd_dense_kernel", by_b, kms.get("detr_criterion_bwd_f32", float("nan")), "one launch, bulk-copy staged")}
    # the reference's path for the same problems on the host cores (all 256 images x 6 layers, a few seconds): per-image
    # Python loop, torch fp32 cost matrix, SciPy-equivalent LSAP (oracle/lsap.c), as detr/matcher.py:40-99 + detr/loss.py:213-217
    from oracle import detr_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sel = list(range(B))
    c_lg, c_bx = logits[sel].cpu(), boxes[sel].cpu()
    c_lab, c_gt = [labels[i].cpu() for i in sel], [gts[i].cpu() for i in sel]
    O.hungarian_match(c_lg[:4, 0], c_bx[:4, 0], c_lab[:4], c_gt[:4], 1.0, 5.0, 2.0)
    t0 = time.perf_counter()
    for l in range(L):
        O.hungarian_match(c_lg[:, l], c_bx[:, l], c_lab, c_gt, 1.0, 5.0, 2.0)
    cpu_s = time.perf_counter() - t0
    return {"roofline_matcher": roof_m, "roofline_criterion": roof_c,
            "metric": "matcher_images_per_sec", "value": round(B / ms_match * 1e3, 1), "unit": "images/s", "ms": round(ms_match, 4),
            "problems": B * L, "sum_gt": sum(counts),
            "workload": "HungarianMatcher only: batch 256 x 6 layers, 100 queries x 1-100 GT boxes, 92 logits, fp32 (BASELINE config 3)",
            "matcher_plus_criterion_fwd_bwd": {"value": round(B / ms_full * 1e3, 1), "unit": "images/s", "ms": round(ms_full, 4)},
            "cpu_baseline": {"value": round(len(sel) / cpu_s, 1), "unit": "images/s", "cores": os.cpu_count() or 1, "kind": "port",
                             "sample": f"{len(sel)} of the 256 images x 6 layers through the oracle port of HungarianMatcher.forward "
                                       f"(per-image loop, torch fp32 cost matrix, single-threaded LSAP), {cpu_s:.2f} s"},
            "note": "assignments are bit-exact vs SciPy on the kernel's costs (tests/test_gpu_matcher.py); the assignment phase is "
                    "latency-bound (serial augmenting paths), not HBM-bound"}


# ------------------------------------------------------------------------------------------------ CPU oracle port
def build_cpu_reference():
    """The oracle port of the reference's path on the host: DetrHarness parameters + oracle encoder/decoder/criterion."""
    from detr_b200.harness import DetrHarness
    from detr_b200.model import DETRConfig
    from oracle import detr_oracle as O

    cfg = DETRConfig(num_classes=NUM_CLASSES)

    def enc_fn(enc, x, pos, mask):
        return O.encoder(dict(enc.named_parameters()), x, pos, mask, cfg.num_encoder_layers, cfg.num_attention_heads, cfg.layer_norm_eps)

    def dec_fn(dec, mem, pos, qe, mask):
        return O.decoder(dict(dec.named_parameters()), mem, pos, qe, mask, cfg.num_decoder_layers, cfg.num_attention_heads, cfg.layer_norm_eps)

    def posenc_fn(eh, ew, heights, widths, scale, feats, temperature):
        pos = O.positional_encoding(eh, ew, heights, widths, scale, feats, temperature)
        return pos.flatten(2).permute(0, 2, 1), O.padding_mask(eh, ew, heights, widths, scale).flatten(1)

    torch.manual_seed(0)
    model = DetrHarness(cfg, encoder_fn=enc_fn, decoder_fn=dec_fn, posenc_fn=posenc_fn).train()
    model.backbone.fold_bn = False   # the reference's stock path: FrozenBatchNorm2d modules, unfused

    class OracleCriterion(torch.nn.Module):
        def forward(self, outputs, targets):
            return O.set_criterion(outputs, targets, NUM_CLASSES, (1.0, 5.0, 2.0), 0.1, 1.0, 5.0, 2.0)

    return model, OracleCriterion()


def build_real_reference():
    """The UNMODIFIED reference classes from the git-ignored baseline/_ref/ (tools/vendor_reference.py copies /root/reference/detr
    there; it travels to the GPU box with the snapshot): detr.model.DETR + detr.matcher.HungarianMatcher + detr.loss.SetCriterion.
    None when the copy is absent."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import refshim
        if refshim.import_reference() is None:
            return None
        import detr.loss as rloss
        import detr.matcher as rmatcher
        import detr.model as rmodel
    except Exception as ex:          # noqa: BLE001 -- any import problem means "fall back to the port", never a crash of the bench
        print(f"bench.py: reference import failed ({ex}); falling back to the oracle port", file=sys.stderr)
        return None
    torch.manual_seed(0)
    model = rmodel.DETR(rmodel.DETRConfig(num_classes=NUM_CLASSES)).train()
    crit = rloss.SetCriterion(NUM_CLASSES, rmatcher.HungarianMatcher(cost_class=1.0, cost_bbox=5.0, cost_giou=2.0),
                              weight_label_ce=1.0, weight_bbox_l1=5.0, weight_bbox_giou=2.0, eos_coef=0.1).train()
    return model, crit


def cpu_step_fn(batch_size: int):
    """-> (step function, cores, kind): one fp32 training step (detr/train.py:258-267 body) of the reference on the host cores:
    the real classes when baseline/_ref is present (kind "reference"), else the oracle port (kind "port")."""
    from detr_b200.harness import make_optimizer, synthetic_batch, train_step
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    real = build_real_reference()
    kind = "reference" if real is not None else "port"
    model, crit = real if real is not None else build_cpu_reference()
    opt = make_optimizer(model, fused=False)
    batch = synthetic_batch(batch_size, H, W, NUM_CLASSES, 20, seed=7)
    return (lambda: float(train_step(model, crit, opt, batch, autocast_dtype=None))), cores, kind


def cpu_baseline(sample_steps: int = 1, batch_size: int = 2):
    """Bounded CPU sample of the same workload: full fp32 training steps of batch 2 (BASELINE config 1 shape)."""
    step, cores, kind = cpu_step_fn(batch_size)
    step()  # warm-up (allocator, oneDNN primitive caches)
    t0 = time.perf_counter()
    for _ in range(sample_steps):
        step()
    dt = (time.perf_counter() - t0) / sample_steps
    what = ("the reference's own classes (baseline/_ref: detr.model.DETR, HungarianMatcher with SciPy, SetCriterion)" if kind == "reference"
            else "the oracle port (oracle/detr_oracle.py + torchvision ResNet-50)")
    return {"value": round(batch_size / dt, 4), "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{sample_steps} fp32 training step(s) of batch {batch_size} at 800x1066 through {what} on the host cores, "
                      f"{dt:.2f} s/step; the LSAP is single-threaded"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch_size = 2
    step, cores, kind = cpu_step_fn(batch_size)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = round(batch_size * args.steps / dt, 4)
    sample = f"each step = one fp32 training step of batch {batch_size} at 800x1066 (bounded sample of the 8 img/GPU workload)"
    print(json.dumps({
        "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": WORKLOAD, "sample": sample, "parallelism": "cpu"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--eager", action="store_true", help="do not capture the step in CUDA graphs (DDP eager path)")
    ap.add_argument("--ncu-step", action="store_true", help="warm up, then run exactly one step between cudaProfilerStart/Stop")
    ap.add_argument("--config", type=int, choices=[2, 4, 5], default=2, help="BASELINE.json config: 2 (default, the headline metric), 4 (DC5 transformer), 5 (R101, 300 queries)")
    ap.add_argument("--graph-profile", action="store_true", help="--config 4: per-call timings inside a CUDA-graph replay (event-record nodes)")
    ap.add_argument("--grid", default="50x67", help="--config 4 only: feature-map size H'xW' (50x67 = DC5 stride 16; 25x34 = the stride-32 map)")
    ap.add_argument("--accum", type=int, default=1, help="gradient-accumulation micro-batches per optimizer step (detr/train.py:116; fixed global batch 64 = 8 img x N GPUs x accum)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
