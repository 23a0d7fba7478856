#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: stall-reason totals and the hottest SASS
instructions (by warp-stall samples).  usage: ncu_src_summary.py file.csv [top]"""
import csv
import sys
import collections

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[2:]:          # first launch only (the export repeats the header per launch)
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr) and r[0] != "Address":
        body.append(r)
tot = sum(int(r[ix["# Samples"]]) for r in body)
inst = sum(int(r[ix["Instructions Executed"]]) for r in body)
print(rows[0][1][:110])
print(f"samples {tot}  warp-instructions {inst}  sass lines {len(body)}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]]) for r in body) for s in stalls}
print("stall totals:", ", ".join(f"{k[6:]}={v * 100 / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 / tot > 0.5))
ops = collections.Counter()
for r in body:
    toks = r[ix["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    ops[op.split(".")[0]] += int(r[ix["Instructions Executed"]])
print("instr mix:", ", ".join(f"{k}={v * 100 / inst:.1f}%" for k, v in ops.most_common(18)))
print(f"{'samples%':>8} {'exec':>10}  sass")
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
    dom = max(stalls, key=lambda s: int(r[ix[s]]))
    print(f"{int(r[ix['# Samples']]) * 100 / tot:8.2f} {int(r[ix['Instructions Executed']]):>10}  {r[ix['Source']].strip()[:90]}   [{dom[6:]}]")
