#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel-name totals and a coarse class split.
usage: python tools/launch_summary.py launches.csv [top_n]"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
cls = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0


def klass(n):
    if re.search(r"detr::|bwd::", n):
        return "own (libdetr_b200)"
    if re.search(r"implicit_gemm|xmma|cudnn|conv|nhwc|max_pool|Padding", n):
        return "backbone conv/pool (cuDNN/ATen)"
    if re.search(r"nvjet|gemm|cublas|splitK", n, re.I):
        return "GEMM (cuBLASLt)"
    if re.search(r"multi_tensor|Optimizer|lpnorm|foreach", n, re.I):
        return "optimizer / clip"
    return "ATen elementwise / reduce / other"


for r in rows[h + 1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    v = v / 1e6 if r[iu] in ("ns", "nsecond") else v / 1e3 if r[iu] in ("us", "usecond") else v
    name = re.sub(r"\(.*", "", r[ik])[:100]
    agg[name][0] += 1; agg[name][1] += v
    c = klass(r[ik]); cls[c][0] += 1; cls[c][1] += v
    tot += v
print(f"total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches")
for k, (n, t) in sorted(cls.items(), key=lambda x: -x[1][1]):
    print(f"  {t:8.3f} ms {100 * t / tot:5.1f}% {n:5d}  {k}")
print()
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{t:8.3f} {100 * t / tot:5.1f}% {n:5d} {k}")
