#!/usr/bin/env python
"""Copy the reference's Python package (/root/reference/detr, read-only, pure Python) into the git-ignored
`baseline/_ref/` so that it travels to the GPU box with the `gpurun` snapshot (it is NOT gpurun-ignored).

    python tools/vendor_reference.py            # no-op (exit 0) when /root/reference is absent, e.g. on the GPU box

Nothing under baseline/_ref is product source: it is the UNMODIFIED reference, used by
  * tests/test_gpu_reference_dropin.py -- patch() applied to the reference's own DETR.forward / train loop,
  * bench.py --impl reference          -- the reference's classes timed on the host cores.
Import it through tests/refshim.py (stubs for the packages the reference imports but this image lacks)."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/detr"
DST = os.path.join(ROOT, "baseline", "_ref", "detr")


def vendor(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not present: nothing to copy ({'found' if os.path.isdir(DST) else 'no'} existing copy at {DST})")
        return os.path.isdir(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    if verbose:
        print(f"copied {SRC} -> {DST} ({len(os.listdir(DST))} files)")
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor() or not os.path.isdir(SRC) else 1)
