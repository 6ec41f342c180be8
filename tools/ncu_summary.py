#!/usr/bin/env python
"""Developer tool: stall reasons, pipe utilisation and DRAM traffic of the first kernel in an ncu report.
usage: python tools/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
vals = rows[2] if len(rows) > 2 else rows[1]
d = dict(zip(hdr, vals))
print("kernel:", d.get("Kernel Name"))
for k, v in d.items():
    if "stalled" in k and k.endswith(".ratio"):
        print("stall %-28s %s" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
for k in ["smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "gpu__time_duration.sum",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]:
    print(k, d.get(k))
