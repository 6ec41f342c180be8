#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > x.csv; python src_hotspots.py x.csv [envs]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
envs = float(sys.argv[2]) if len(sys.argv) > 2 else 1048576.0
fname = None; hdr = None
agg = defaultdict(lambda: [0, 0, ""])
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iX = hdr.index("Instructions Executed"); iS = hdr.index("Warp Stall Sampling (All Samples)"); continue
    if hdr is None or len(r) < len(hdr): continue
    # cuda lines have Line No set and aggregated metrics; sass rows have empty line no
    if r[0].strip() == "": continue
    try:
        x = int(r[iX]); s = int(r[iS])
    except ValueError:
        continue
    key = (fname, int(r[0]))
    agg[key][0] += x; agg[key][1] += s; agg[key][2] = r[1].strip()[:110]
tot = sum(v[0] for v in agg.values()); st = sum(v[1] for v in agg.values())
print("total instr/env %.1f  samples %d" % (tot / envs, st))
for (f, ln), (x, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:70]:
    print("%-22s %5d %8.1f %6.1f%% %6.1f%%  %s" % (f, ln, x / envs, 100.0 * x / tot, 100.0 * s / max(st, 1), src))
