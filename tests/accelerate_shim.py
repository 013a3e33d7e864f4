"""A local stand-in for the ~16 members of HuggingFace Accelerate that the reference's training script touches
(detr/train.py:13-14,106-124,218-220,258-267,271-286,320; SURVEY.md 8b "Caveat"): `accelerate` is not installed in this image
and there is no network, so this is what makes "the B200 classes drop into detr/train.py unchanged" testable.  Test
infrastructure only; it mirrors Accelerate's documented single-process / DDP behaviour:

  * prepare(): modules to the device (DistributedDataParallel when torch.distributed has > 1 rank), model.forward wrapped in
    autocast with fp32-converted outputs under mixed precision, data loaders yield batches moved to the device, the optimizer
    skips step() / zero_grad() on non-boundary micro-steps;
  * accumulate(model): gradient synchronisation (and the optimizer step) only every `gradient_accumulation_steps`-th
    micro-step -- DDP.no_sync() in between; backward(loss) divides by the accumulation count; clip_grad_norm_ only on
    boundary steps;
  * trackers / checkpoints: logs are appended to `Accelerator.logs`; save_state() writes model.safetensors per model."""
from __future__ import annotations

import contextlib
import os
import types
from dataclasses import dataclass
from typing import Any, Optional

import torch
import torch.distributed as dist


@dataclass
class ProjectConfiguration:
    project_dir: Optional[str] = None
    logging_dir: Optional[str] = None
    automatic_checkpoint_naming: bool = False
    total_limit: Optional[int] = None
    save_on_each_node: bool = False
    iteration: int = 0


def gather_object(obj: Any):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, obj)
        return [x for part in out for x in (part if isinstance(part, list) else [part])]
    return obj


def _to_device(x, device):
    if torch.is_tensor(x):
        return x.to(device, non_blocking=True)
    if isinstance(x, dict):
        return {k: _to_device(v, device) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_to_device(v, device) for v in x)
    return x


def _to_fp32(x):
    if torch.is_tensor(x):
        return x.float() if x.is_floating_point() else x
    if isinstance(x, dict):
        return {k: _to_fp32(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_to_fp32(v) for v in x)
    return x


class _Loader:
    def __init__(self, loader, device):
        self._loader, self._device = loader, device
        self.dataset = getattr(loader, "dataset", None)

    def __len__(self):
        return len(self._loader)

    def __iter__(self):
        for batch in self._loader:
            yield _to_device(batch, self._device)


class _Optimizer:
    """optimizer.step() / zero_grad() take effect only on gradient-synchronisation steps (Accelerate's AcceleratedOptimizer)."""

    def __init__(self, optimizer, accelerator):
        self.optimizer, self._acc = optimizer, accelerator

    def __getattr__(self, name):
        return getattr(self.optimizer, name)

    @property
    def param_groups(self):
        return self.optimizer.param_groups

    def step(self, *a, **k):
        if self._acc.sync_gradients:
            return self.optimizer.step(*a, **k)

    def zero_grad(self, *a, **k):
        if self._acc.sync_gradients:
            return self.optimizer.zero_grad(*a, **k)


class _Tracker:
    def __init__(self):
        self.writer = types.SimpleNamespace(add_image=lambda *a, **k: None, add_images=lambda *a, **k: None)


class Accelerator:
    def __init__(self, mixed_precision: str = "no", log_with=None, project_config: Optional[ProjectConfiguration] = None,
                 step_scheduler_with_optimizer: bool = True, split_batches: bool = False, gradient_accumulation_steps: int = 1,
                 device: Optional[torch.device] = None, **_):
        self.mixed_precision = mixed_precision
        self.project_config = project_config or ProjectConfiguration()
        self.gradient_accumulation_steps = max(int(gradient_accumulation_steps), 1)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        if device is None:
            device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))) if torch.cuda.is_available() else torch.device("cpu")
        self.device = device
        self.sync_gradients = True
        self._micro_step = 0
        self._models, self.logs = [], []
        self._tracker = _Tracker()

    # -- process info / logging -----------------------------------------------------------------------
    @property
    def is_main_process(self):
        return self.rank == 0

    @property
    def is_local_main_process(self):
        return int(os.environ.get("LOCAL_RANK", 0)) == 0

    def print(self, *a, **k):
        if self.is_local_main_process:
            print(*a, **k)

    def init_trackers(self, name, config=None):
        self.run_name = name

    def get_tracker(self, name):
        return self._tracker

    def log(self, values, step=None):
        self.logs.append((step, values))

    def end_training(self):
        pass

    # -- prepare ----------------------------------------------------------------------------------------
    def _prepare_one(self, obj):
        if isinstance(obj, torch.nn.Module):
            obj = obj.to(self.device)
            has_params = any(p.requires_grad for p in obj.parameters())
            if self.world > 1 and has_params:
                obj = torch.nn.parallel.DistributedDataParallel(obj, device_ids=[self.device.index] if self.device.type == "cuda" else None)
            if self.mixed_precision in ("bf16", "fp16") and has_params:
                dtype = torch.bfloat16 if self.mixed_precision == "bf16" else torch.float16
                inner = obj.forward

                def forward(*a, __inner=inner, **k):    # autocast(model.forward) + convert_outputs_to_fp32, as Accelerate.prepare_model
                    with torch.autocast(device_type=self.device.type, dtype=dtype):
                        out = __inner(*a, **k)
                    return _to_fp32(out)

                obj.forward = forward
            if has_params:
                self._models.append(obj)
            return obj
        if isinstance(obj, torch.optim.Optimizer):
            return _Optimizer(obj, self)
        if isinstance(obj, torch.utils.data.DataLoader) or hasattr(obj, "__iter__") and hasattr(obj, "__len__") and not isinstance(obj, (list, tuple, dict)):
            return _Loader(obj, self.device)
        return obj   # LR schedulers step once per epoch in the reference (step_scheduler_with_optimizer=False)

    def prepare(self, *objs):
        out = tuple(self._prepare_one(o) for o in objs)
        return out if len(out) > 1 else out[0]

    # -- step -------------------------------------------------------------------------------------------
    @contextlib.contextmanager
    def accumulate(self, *models):
        self._micro_step += 1
        self.sync_gradients = self._micro_step % self.gradient_accumulation_steps == 0
        ddp = [m for m in models if isinstance(m, torch.nn.parallel.DistributedDataParallel)]
        with contextlib.ExitStack() as stack:
            if not self.sync_gradients:
                for m in ddp:
                    stack.enter_context(m.no_sync())
            yield

    @contextlib.contextmanager
    def autocast(self):
        if self.mixed_precision in ("bf16", "fp16"):
            with torch.autocast(device_type=self.device.type, dtype=torch.bfloat16 if self.mixed_precision == "bf16" else torch.float16):
                yield
        else:
            yield

    def backward(self, loss, **k):
        (loss / self.gradient_accumulation_steps).backward(**k)

    def clip_grad_norm_(self, parameters, max_norm, norm_type=2):
        if self.sync_gradients:
            return torch.nn.utils.clip_grad_norm_(parameters, max_norm, norm_type=norm_type)
        return None

    def save_state(self, output_dir: Optional[str] = None):
        from safetensors.torch import save_model
        d = output_dir or os.path.join(self.project_config.project_dir or ".", "checkpoints", f"checkpoint_{self.project_config.iteration}")
        if self.is_main_process:
            os.makedirs(d, exist_ok=True)
            for i, m in enumerate(self._models):
                inner = m.module if hasattr(m, "module") else m
                save_model(inner, os.path.join(d, "model.safetensors" if i == 0 else f"model_{i}.safetensors"))
        self.project_config.iteration += 1
        return d


utils = types.ModuleType("accelerate.utils")
utils.ProjectConfiguration = ProjectConfiguration
utils.gather_object = gather_object
