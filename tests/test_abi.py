"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/detr_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "detr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|int64_t)\s+(detr_\w+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _declared()
    assert "detr_hungarian_match_f32" in names and "detr_criterion_fwd_f32" in names and len(names) >= 10


def test_library_exports_every_declared_symbol():
    from detr_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run `python detr-object-detection_b200/build.py` (or __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/detr_b200.h but not exported"
    assert lib.detr_b200_abi_version() >= 1


def test_python_binding_covers_header():
    from detr_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    _lib.load()


def test_no_cpu_fallback():
    import torch
    from detr_b200 import HungarianMatcher
    m = HungarianMatcher(1, 5, 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(1, 4, 3), torch.rand(1, 4, 4), [torch.zeros(1, dtype=torch.int64)], [torch.tensor([[0.1, 0.1, 0.5, 0.5]])])
