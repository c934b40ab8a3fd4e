#!/usr/bin/env python3
"""Hot SASS regions of an ncu report: python tools/ncu_hot.py gpurun_out/prof_X.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; data = rows[hi+1:]
si = h.index("Warp Stall Sampling (All Samples)"); ii = h.index("Instructions Executed"); src = h.index("Source")
tot = sum(int(r[si] or 0) for r in data); toti = sum(int(r[ii] or 0) for r in data)
print(f"total samples {tot}, instructions {toti}, sass lines {len(data)}")
idx = sorted(range(len(data)), key=lambda i: -int(data[i][si] or 0))[:top]
for i in sorted(idx):
    r = data[i]
    print(f"{i:5d} {100*int(r[si] or 0)/tot:5.1f}%  exec {int(r[ii] or 0):9d}  {r[src][:110]}")
