"""Builds libdetr_b200.so (all CUDA kernels + the C ABI of include/detr_b200.h) for sm_100a, in-tree.

    python detr-object-detection_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "detr_b200", "libdetr_b200.so")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# development switches, e.g. DETR_B200_DEFINES="-DDETR_BWD_TIMELINE" (clock64 stamps for tools/attn_timeline.py)
FLAGS += [f for f in os.environ.get("DETR_B200_DEFINES", "").split() if f]


def _stamp() -> str:
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for f in sorted(glob.glob(os.path.join(CSRC, "*")) + [os.path.join(HERE, "..", "include", "detr_b200.h")]):
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode() + fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp_file = os.path.join(OBJ, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(OUT) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return OUT
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "detr_b200.h")]
    newest_h = max(os.path.getmtime(h) for h in headers)
    flags_file = os.path.join(OBJ, "flags")
    same_flags = os.path.exists(flags_file) and open(flags_file).read() == " ".join(FLAGS)

    def cc(pair):
        s, o = pair
        if not force and os.path.exists(o) and os.path.getmtime(o) > max(os.path.getmtime(s), newest_h) and same_flags:
            return
        cmd = [NVCC, *FLAGS, "-c", s, "-o", o] + (["-Xptxas", "-v"] if verbose else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {s}")

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        list(ex.map(cc, zip(srcs, objs)))
    subprocess.check_call([NVCC, "-shared", "-o", OUT, *objs, "-lcudart", "-lcuda"])
    with open(stamp_file, "w") as f:
        f.write(stamp)
    with open(flags_file, "w") as f:
        f.write(" ".join(FLAGS))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
