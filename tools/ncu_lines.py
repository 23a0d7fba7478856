#!/usr/bin/env python
"""Per-source-line warp-instruction counts from `ncu --page source --csv --print-source cuda,sass`.
usage: ncu_lines.py file.csv [norm] [top]   (norm: divide counts, e.g. by the warps of particles)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
norm = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
cur = hdr = None
out = []
for r in rows:
    if r and r[0] == "File Path":
        cur, hdr = r[1], None
        continue
    if r and r[0] == "Line No":
        hdr = r
        ie_i = hdr.index("Instructions Executed")
        sm_i = hdr.index("# Samples")
        continue
    if hdr and cur and len(r) == len(hdr) and r[0] != "":
        try:
            ie = int(r[ie_i])
        except ValueError:
            ie = 0
        if ie > 0:
            out.append((ie, cur.split("/")[-1], r[0], r[1].strip()[:110], r[sm_i]))
tot = sum(o[0] for o in out)
print("total warp-instructions", tot, " normalised", round(tot / norm, 1))
for o in sorted(out, key=lambda x: -x[0])[:top]:
    print(f"{o[0] / norm:8.1f} {o[4]:>7} {o[1]}:{o[2]}  {o[3]}")
