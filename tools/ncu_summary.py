#!/usr/bin/env python3
"""Summarise gpurun_out/launches.csv and gpurun_out/prof_*.ncu-rep (read here with ncu -i, no GPU needed)."""
import csv, glob, os, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("PDDP_NCU_DIR") or os.path.join(ROOT, "gpurun_out")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
p = os.path.join(OUT, "launches.csv")
if os.path.exists(p):
    rows = [r for r in csv.reader(open(p)) if len(r) > 5]
    h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
    t = defaultdict(list)
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", "")); v = v / 1000 if r[ui] == "ns" else v
            t[r[ki].split("(")[0][:44]].append(v)
        except ValueError:
            pass
    print("== launch list (us, ncu-serialised)")
    tot = sum(sum(v) for v in t.values())
    for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {k:46s} n={len(v):3d} avg {sum(v)/len(v):9.1f} us   share {100*sum(v)/tot:5.1f}%")
for rep in sorted(glob.glob(os.path.join(OUT, "prof_*.ncu-rep"))):
    if len(sys.argv) > 1 and not any(a in rep for a in sys.argv[1:]):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        continue
    h, units, v = rows[0], rows[1], rows[-1]
    print("==", os.path.basename(rep), v[h.index("Kernel Name")][:60] if "Kernel Name" in h else "")
    for w in WANT:
        if w in h:
            print(f"  {w:88s} {v[h.index(w)]:>16s} {units[h.index(w)]}")
