#!/usr/bin/env python
"""Summarise an .ncu-rep (run here, no GPU needed): per-launch headline metrics + instruction mix + top stall lines.
usage: python tools/ncu_summary.py report.ncu-rep [kernel-regex] [launch-index]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "."
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_xu.sum", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
        "smsp__inst_executed_pipe_xu.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
ik = hdr.index("Kernel Name")
for n, r in enumerate(rows[2:]):
    if not re.search(kre, r[ik]):
        continue
    print(f"## launch {n}: {r[ik][:90]}")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w}: {r[i]} {units[i]}")
if len(sys.argv) > 3:
    li = sys.argv[3]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", li, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = 0 if "Source" in rows[0] else 1
    h = rows[hi]
    iS, iN, iP = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    data = [(r[iS].strip(), int(r[iN]), int(r[iP])) for r in rows[hi + 1:] if len(r) > iN and r[iN].isdigit()]
    tot, ts = sum(d[1] for d in data), sum(d[2] for d in data)
    print(f"\n## source page launch {li}: {tot} warp-instructions, {ts} samples, {len(data)} SASS lines")
    op, ops = collections.Counter(), collections.Counter()
    for s, n, sm in data:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", s)
        k = m.group(2) if m else s[:10]
        op[k] += n; ops[k] += sm
    for k, v in op.most_common(22):
        print(f"  {k:10s} {v:10d} {100*v/tot:5.1f}%  stall-samples {100*ops[k]/max(ts,1):5.1f}%")
    print("  -- top sampled lines")
    for s, n, sm in sorted(data, key=lambda d: -d[2])[:14]:
        print(f"  {sm:5d} {n:9d} {s[:90]}")
