#!/usr/bin/env python
"""Turn a gpurun_out/<tag>/ directory (launches.csv + prof_*.ncu-rep from tools/gpu_check.sh) into
the text summaries committed under profiles/.   usage: ncu_summarize.py gpurun_out/r01a profiles/r01a"""
import collections
import csv
import os
import re
import subprocess
import sys

src, dst = sys.argv[1], sys.argv[2]
out = []

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
           "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "smsp__inst_executed.sum"]

lp = os.path.join(src, "launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 10]
    ix = {h: i for i, h in enumerate(rows[0])}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])[:70]
        v, u = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[u]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out.append("## launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n")
    out.append(f"{'kernel':72s} {'n':>4s} {'total ms':>10s} {'avg us':>10s} {'share':>6s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        out.append(f"{k:72s} {n:4d} {t / 1e3:10.3f} {t / n:10.1f} {t / tot * 100:5.1f}%")
    out.append("")

for f in sorted(os.listdir(src)):
    # either the report itself or its `--page raw --csv` export made on the GPU box (reports are ~40 MB each and
    # gpurun brings back at most 64 MiB per call)
    if f.endswith("_raw.csv"):
        raw = open(os.path.join(src, f)).read()
    elif f.endswith(".ncu-rep") and not os.path.exists(os.path.join(src, f[:-8] + "_raw.csv")):
        raw = subprocess.run(["ncu", "-i", os.path.join(src, f), "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
    else:
        continue
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out.append(f"## {f}: ncu --set full --clock-control none (per launch)\n")
    for r in rows[2:]:
        out.append("### " + re.sub(r"\(.*", "", r[ix["Kernel Name"]]))
        for m in METRICS:
            if m in ix:
                out.append(f"  {m:70s} {r[ix[m]]:>16s} {units[ix[m]]}")
        out.append("")
os.makedirs(os.path.dirname(dst) or ".", exist_ok=True)
open(dst + "_ncu.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
