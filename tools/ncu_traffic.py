#!/usr/bin/env python
"""Turn committed ncu captures into the numbers bench.py reports (profiles/r02_ncu_numbers.json), so that no measured value is a
literal in bench.py:

    python tools/ncu_traffic.py --rep gpurun_out/<capture>.ncu-rep [--launches profiles/<step launch list>.csv]

From the `--set full` capture (tools/kernel_bench.py --only attention,matcher --iters 1 under ncu): per-launch
dram__bytes_read.sum + dram__bytes_write.sum, grouped into the calls bench.py names.  From the step launch list
(`bench.py --ncu-step` under `ncu --metrics gpu__time_duration.sum`): the share of libdetr_b200 kernels in the step."""
import argparse
import csv
import io
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_ncu_numbers.json")


def raw_rows(rep):
    """`rep`: an .ncu-rep, or the text of `ncu -i <rep> --page raw --csv` made on the GPU box (the report itself can exceed what
    travels back)."""
    txt = open(rep, errors="replace").read() if rep.endswith(".csv") else \
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    txt = txt[txt.index('"ID"'):]
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    it = hdr.index("gpu__time_duration.sum")
    ig = hdr.index("launch__grid_size")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = []
    for r in rows[2:]:
        out.append({"name": r[ik], "bytes": float(r[ir].replace(",", "")) * scale.get(units[ir], 1.0) + float(r[iw].replace(",", "")) * scale.get(units[iw], 1.0),
                    "us": float(r[it].replace(",", "")) * (1e-3 if units[it] in ("ns", "nsecond") else 1.0), "grid": int(r[ig].replace(",", ""))})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rep", default="")
    ap.add_argument("--launches", default="")
    args = ap.parse_args()
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    if args.rep:
        rows = raw_rows(args.rep)
        res["source_capture"] = os.path.basename(args.rep)
        res["launches"] = [{"kernel": re.sub(r"\(.*", "", r["name"])[:80], "grid": r["grid"], "us": round(r["us"], 2), "dram_bytes": round(r["bytes"])} for r in rows]
        # kernel_bench --only attention runs, per shape, forward (main [+ combine]) then backward (delta, main, dq reduce [, dkv reduce]); the
        # first backward main kernel belongs to the encoder shape (config 2), the fourth to DC5 (config 4)
        bwd = [i for i, r in enumerate(rows) if "attention_bwd_kernel" in r["name"]]

        def call_bytes(i):
            tot, j = rows[i]["bytes"], i - 1
            if j >= 0 and "attention_delta" in rows[j]["name"]:
                tot += rows[j]["bytes"]
            j = i + 1
            while j < len(rows) and re.search(r"attention_dq_reduce|attention_dkv_reduce", rows[j]["name"]):
                tot += rows[j]["bytes"]
                j += 1
            return round(tot)
        if bwd:
            res["attention_bwd_encoder"] = call_bytes(bwd[0])
        if len(bwd) >= 4:
            res["attention_bwd_dc5"] = call_bytes(bwd[3])
        for key, pat in (("hungarian_match_kernel", "hungarian_match_kernel"), ("criterion_fwd", "criterion_fwd_dense_kernel"),
                         ("criterion_bwd_dense_kernel", "criterion_bwd_dense_kernel")):
            hit = [r for r in rows if pat in r["name"]]
            if hit:
                res[key] = round(hit[0]["bytes"])
    if args.launches:
        rows = list(csv.reader(open(args.launches, errors="replace")))
        h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        hdr = rows[h]
        ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        tot = own = 0.0
        for r in rows[h + 1:]:
            if len(r) > iv:
                v = float(r[iv].replace(",", ""))
                tot += v
                if re.search(r"detr::|bwd::", r[ik]):
                    own += v
        res["own_share_of_step"] = round(own / tot, 4)
        res["source_launch_list"] = os.path.basename(args.launches)
    json.dump(res, open(OUT, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "launches"}, indent=1))


if __name__ == "__main__":
    main()
