"""ctypes binding of libdetr_b200.so (C ABI declared in include/detr_b200.h).

The library is the product: there is NO fallback.  If the shared object is missing or the device is not an
sm_100 part, importing a kernel entry point raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int32, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdetr_b200.so")
_lib = None
_checked_devices: set[int] = set()

P = c_void_p
_STRIDES3 = [c_int64, c_int64, c_int64]

# name -> argtypes ; mirrors include/detr_b200.h one to one
SIGNATURES = {
    "detr_b200_abi_version": [],
    "detr_b200_last_error": [ctypes.c_char_p, c_int],
    "detr_b200_check_device": [c_int],
    "detr_matcher_smem_bytes": [c_int, c_int, c_int],
    "detr_cost_matrix_f32": [P, *_STRIDES3, P, *_STRIDES3, P, P, P, c_int, c_int, c_int, c_int, c_int,
                             c_float, c_float, c_float, P, P, P],
    "detr_hungarian_match_f32": [P, *_STRIDES3, P, *_STRIDES3, P, P, P, P, c_int, c_int, c_int, c_int, c_int,
                                 c_float, c_float, c_float, P, P, P, P, P, P],
    "detr_lsap_f32": [P, P, P, P, c_int, c_int, c_int, P, P, P, P, P],
    "detr_lsap_f64": [P, P, P, P, c_int, c_int, c_int, P, P, P, P, P],
    "detr_criterion_fwd_f32": [P, *_STRIDES3, P, *_STRIDES3, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int,
                               c_float, c_float, c_float, P, P, P, P, P, P, P, P],
    "detr_criterion_bwd_f32": [P, P, *_STRIDES3, P, *_STRIDES3, P, P, P, P, P, P, P,
                               c_int, c_int, c_int, c_int, c_float, c_float, c_float, P, P, P],
    "detr_attention_fwd_workspace_floats": [c_int, c_int, c_int, c_int],
    "detr_attention_fwd_bf16": [P, c_int64, c_int64, P, c_int64, c_int64, P, c_int64, c_int64, P, c_int64, c_int64, P,
                                P, P, P, c_int64, P, c_int, c_int, c_int, c_int, c_float, ctypes.c_uint64, P, P],
    "detr_attention_bwd_workspace_floats": [c_int, c_int, c_int, c_int],
    "detr_attention_bwd_bf16": [P, c_int64, c_int64] * 4 + [P] + [P, c_int64, c_int64] + [P, P, P] + [P, c_int64, c_int64] * 3 +
                               [P, c_int64, P, c_int, c_int, c_int, c_int, c_float, ctypes.c_uint64, P, P],
    "detr_colsum_chunks": [c_int, c_int],
    "detr_colsum_bf16": [P, c_int64, c_int, c_int, P, P, P, P],
    "detr_layernorm_grid": [c_int],
    "detr_layernorm_fwd": [P, c_int, c_int64, P, P, P, c_int, c_int64, c_int64, c_int, P, P, c_int, P, P, c_int, c_int, c_float, P],
    "detr_layernorm_bwd": [P, P, c_int, P, P, c_int, c_int64, P, P, P, P, P, P, P, P, c_int, c_int, P],
    "detr_layernorm_bwd_tail": [P, P, c_int, P, P, c_int, c_int64, P, P, P, P, P, P, P, P, c_int, c_int, P, P, c_float, ctypes.c_uint64, P, P],
    "detr_layernorm_bwd_fold": [P, c_int, c_int, P, P, P, P],
    "detr_gemm_ln_partition": [c_int, c_int, c_int, c_int, P, P],
    "detr_heads_grad_prep": [P, c_int, P, P, P, c_int, P, c_int, c_int, P],
    "detr_epilogue_fwd": [c_int, P, c_int, P, P, c_int, c_int, c_float, ctypes.c_uint64, P, P],
    "detr_epilogue_chunks": [c_int, c_int],
    "detr_scale_cast_multi": [P, c_int, c_int, P],
    "detr_positional_encoding_f32": [P, P, c_int, c_int, c_int, c_int, c_int, c_float, P, P, P],
    "detr_sumsq_grid": [ctypes.c_longlong],
    "detr_sumsq_f32": [P, ctypes.c_longlong, P, P, P, P, P],
    "detr_adamw_clip_f32": [P, P, P, P, ctypes.c_longlong, P, c_float, c_float, c_float, c_float, P, P, c_float, c_float, P],
    "detr_add_relu_mask_bf16": [P, P, P, P, ctypes.c_longlong, P],
    "detr_stem_s2d_bf16": [P, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, c_int, c_int, c_int, c_int, P, c_int, P],
    "detr_maxpool3x3s2_out": [c_int],
    "detr_maxpool3x3s2_fwd_bf16": [P, P, P, c_int, c_int, c_int, c_int, P],
    "detr_maxpool3x3s2_bwd_bf16": [P, P, P, c_int, c_int, c_int, c_int, P],
    "detr_epilogue_bwd": [c_int, P, c_int, P, P, P, P, P, c_int, c_int, c_float, ctypes.c_uint64, P, P],
    "detr_gemm_bf16": [P, c_int64, P, c_int64, c_int, c_int, c_int, c_int, c_int, P, P, c_int, c_int64, P, c_int64, P, c_int64,
                       c_float, ctypes.c_uint64, P, P],
    "detr_gemm_ln_bf16": [P, c_int, c_int64, P, P, c_float, P, c_int64, c_int64, c_int, c_int, P, c_int64, c_int, c_int, c_int, P, P, c_int64,
                          P, c_int64, P, P, P, P, c_float, ctypes.c_uint64, P, P],
    "detr_gemm_wgrad_workspace_floats": [c_int, c_int, c_int],
    "detr_gemm_wgrad_bf16": [P, c_int64, P, c_int64, P, c_int64, c_int, c_int, c_int, c_int, P, P, P, P],
}
_RESTYPE = {"detr_gemm_wgrad_workspace_floats": c_int64, "detr_matcher_smem_bytes": c_int64, "detr_attention_bwd_workspace_floats": c_int64,
            "detr_attention_fwd_workspace_floats": c_int64}


def load() -> ctypes.CDLL:
    """Load the shared library (built by detr-object-detection_b200/build.py). Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python detr-object-detection_b200/build.py` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI drifted
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, c_int)
        _lib = lib
    return _lib


FOLD_MAX_TENSORS = 64


class FoldTable(ctypes.Structure):
    """Mirror of DetrFoldTable (include/detr_b200.h)."""
    _fields_ = [("src", c_void_p * FOLD_MAX_TENSORS), ("dst", c_void_p * FOLD_MAX_TENSORS), ("scale", c_void_p * FOLD_MAX_TENSORS),
                ("O", c_int * FOLD_MAX_TENSORS), ("I", c_int * FOLD_MAX_TENSORS), ("HW", c_int * FOLD_MAX_TENSORS),
                ("src_stride", (c_int * 3) * FOLD_MAX_TENSORS), ("dst_stride", (c_int * 3) * FOLD_MAX_TENSORS), ("n", c_int)]


# kernels launched per C-ABI call (bench.py's gpu_launches is counted from this table; callers whose launch count depends on
# the shape pass the exact number through call(..., launches=))
KERNELS_PER_CALL = {"detr_cost_matrix_f32": 1, "detr_hungarian_match_f32": 1, "detr_lsap_f32": 1, "detr_lsap_f64": 1,
                    "detr_criterion_fwd_f32": 3, "detr_criterion_bwd_f32": 1, "detr_attention_fwd_bf16": 2,
                    "detr_attention_bwd_bf16": 4, "detr_colsum_bf16": 1, "detr_layernorm_fwd": 1, "detr_layernorm_bwd": 2, "detr_layernorm_bwd_tail": 2, "detr_layernorm_bwd_fold": 1, "detr_heads_grad_prep": 1,
                    "detr_epilogue_fwd": 1, "detr_epilogue_bwd": 1, "detr_scale_cast_multi": 1,
                    "detr_maxpool3x3s2_fwd_bf16": 1, "detr_maxpool3x3s2_bwd_bf16": 1,
                    "detr_gemm_bf16": 1, "detr_gemm_ln_bf16": 1, "detr_gemm_wgrad_bf16": 2,
                    "detr_positional_encoding_f32": 1, "detr_sumsq_f32": 1, "detr_adamw_clip_f32": 1, "detr_add_relu_mask_bf16": 1, "detr_stem_s2d_bf16": 1}
launch_count = 0          # kernels of libdetr_b200.so launched by this process
_profile = None           # when a list: (name, tag, start_event, end_event) per call
_profile_external = False


class profile:
    """Context manager: CUDA-event timing of every C-ABI launch on its own stream.

        with _lib.profile() as prof: step()
        torch.cuda.synchronize(); rows = prof.summary()   # {(name, tag): (calls, total_ms)}
    """

    def __init__(self, external: bool = False):
        # external=True: the events become timable event-record NODES when the calls are made under CUDA-graph capture -- per-call
        # durations INSIDE a graph replay (read them after a replay + synchronize)
        self.external = external

    def __enter__(self):
        global _profile, _profile_external
        self.records = []
        _profile = self.records
        _profile_external = self.external
        return self

    def __exit__(self, *exc):
        global _profile
        _profile = None
        return False

    def summary(self):
        out = {}
        for name, tag, a, b in self.records:
            n, t = out.get((name, tag), (0, 0.0))
            out[(name, tag)] = (n + 1, t + a.elapsed_time(b))
        return out

    def median(self):
        """{(name, tag): median ms per call}"""
        per = {}
        for name, tag, a, b in self.records:
            per.setdefault((name, tag), []).append(a.elapsed_time(b))
        return {k: sorted(v)[len(v) // 2] for k, v in per.items()}


_NUM_SMS = {}


def num_sms() -> int:
    """SM count of the current device (the persistent kernels launch min(#SMs, #items) CTAs)."""
    dev = torch.cuda.current_device()
    if dev not in _NUM_SMS:
        _NUM_SMS[dev] = torch.cuda.get_device_properties(dev).multi_processor_count
    return _NUM_SMS[dev]


def call(name: str, *args, tag=None, launches=None) -> None:
    """Invoke a launcher of the C ABI, count its kernels, raise on a non-zero return code."""
    global launch_count
    fn = getattr(load(), name)
    launch_count += KERNELS_PER_CALL[name] if launches is None else launches
    if _profile is None:
        rc = fn(*args)
    else:
        kw = {"external": True} if _profile_external else {}
        a, b = torch.cuda.Event(enable_timing=True, **kw), torch.cuda.Event(enable_timing=True, **kw)
        a.record()
        rc = fn(*args)
        b.record()
        _profile.append((name, tag, a, b))
    check(rc, name)


def last_error() -> str:
    buf = ctypes.create_string_buffer(512)
    load().detr_b200_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = last_error()
        if "all costs can't be 0" in msg:
            raise AssertionError(msg)  # detr/matcher.py:38
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor, got {t.device}; detr_b200 has no CPU path")
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if dev not in _checked_devices:
        check(load().detr_b200_check_device(dev), "device check")
        _checked_devices.add(dev)


_zero_counters: dict = {}


def zero_counters(device: torch.device) -> torch.Tensor:
    """Per-device block of 64 zeroed uint32 used by single-launch reductions (each kernel leaves it zeroed)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    t = _zero_counters.get(key)
    if t is None:
        t = torch.zeros(64, dtype=torch.int32, device=device)
        _zero_counters[key] = t
    return t


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()
