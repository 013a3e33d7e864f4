"""detr_b200 -- B200-native (sm_100a) implementation of the DETR data-parallel training hot path of
anenbergb/DETR-object-detection, behind the reference's own class signatures.

Host code is Python/PyTorch; every kernel lives in libdetr_b200.so (C ABI: include/detr_b200.h).
"""
from . import _lib
from .loss import SetCriterion
from .matcher import HungarianMatcher, linear_sum_assignment_cuda
from .targets import PackedTargets, StaticTargets, pack_targets

__all__ = ["HungarianMatcher", "SetCriterion", "linear_sum_assignment_cuda", "PackedTargets", "StaticTargets", "pack_targets", "_lib"]
