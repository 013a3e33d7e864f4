#!/usr/bin/env python
"""Instruction mix and hottest SASS lines of one launch in an .ncu-rep (source page; needs --import-source on).
usage: python tools/ncu_source.py report.ncu-rep launch_index [top_lines]"""
import collections, csv, io, re, subprocess, sys

rep, li = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", li, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
iS, iN, iP = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
data = []
for r in rows:
    if len(r) > iN and r[iN].isdigit():
        data.append((r[iS].strip(), int(r[iN]), int(r[iP]) if r[iP].isdigit() else 0))
tot, ts = sum(d[1] for d in data), sum(d[2] for d in data)
print(f"{tot} warp-instructions, {ts} samples, {len(data)} SASS lines")
op, ops = collections.Counter(), collections.Counter()
for s, n, sm in data:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", s)
    k = m.group(2) if m else s[:10]
    op[k] += n; ops[k] += sm
for k, v in op.most_common(28):
    print(f"  {k:10s} {v:10d} {100 * v / tot:5.1f}%  stall-samples {100 * ops[k] / max(ts, 1):5.1f}%")
print("hottest lines by stall samples:")
for s, n, sm in sorted(data, key=lambda d: -d[2])[:top]:
    print(f"  {sm:6d} {n:9d}  {s[:110]}")
