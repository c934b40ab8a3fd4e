#!/usr/bin/env python3
"""Executed warp-instructions and stall samples aggregated by CUDA source line: python tools/ncu_lines.py rep [top]"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = ""
agg = defaultdict(lambda: [0, 0, ""])
h = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]; h = None; continue
    if r and r[0] == "Line No":
        h = r; continue
    if h and len(r) == len(h) and "Instructions Executed" in h:
        try:
            ln = int(r[0])
        except ValueError:
            continue
        ie = int(r[h.index("Instructions Executed")] or 0); st = int(r[h.index("Warp Stall Sampling (All Samples)")] or 0)
        key = (fname, ln)
        agg[key][0] += ie; agg[key][1] += st
        if not agg[key][2]:
            agg[key][2] = r[1]
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f"total instr {tot_i}  samples {tot_s}")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/max(tot_i,1):5.1f}% instr {100*v[1]/max(tot_s,1):5.1f}% stall  {f}:{ln:<4d} {v[2].strip()[:100]}")
