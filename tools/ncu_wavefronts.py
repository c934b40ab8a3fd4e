#!/usr/bin/env python3
"""Shared-memory wavefronts / executed instructions per CUDA source line of one kernel: joins the SASS rows of an ncu report
(--page source --csv, address order) with nvdisasm -g line info of the same kernel in libpddp.so (same instruction order).
usage: tools/ncu_wavefronts.py <report.ncu-rep> <kernel-substring-mangled> [top] [per_unit_divisor]"""
import collections, csv, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40; div = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); h = rows[hi]; data = rows[hi+1:]
d = tempfile.mkdtemp(); subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "parallel-ddp_b200", "libpddp.so")], cwd=d, capture_output=True)
lines = []
for f in os.listdir(d):
    s = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
    cur = None; line = None
    for ln in s.splitlines():
        if ln.startswith(".text."): cur = ln; line = None; continue
        if cur is None or kern not in cur: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): lines.append(line)
    if lines: break
assert len(lines) == len(data), (len(lines), len(data))
ie = h.index("Instructions Executed"); wf = h.index("L1 Wavefronts Shared"); wi = h.index("L1 Wavefronts Shared Ideal"); st = h.index("Warp Stall Sampling (All Samples)")
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for l, r in zip(lines, data):
    a = agg[l]; a[0] += int(r[ie] or 0); a[1] += int(r[wf] or 0); a[2] += int(r[wi] or 0); a[3] += int(r[st] or 0)
T = [sum(a[i] for a in agg.values()) for i in range(4)]
print(f"total: instr {T[0]/div:.1f}  wavefronts {T[1]/div:.1f} (ideal {T[2]/div:.1f})  samples {T[3]}")
for l, a in sorted(agg.items(), key=lambda kv: -(kv[1][1]*2 + kv[1][0]))[:top]:
    print(f"{(l[0]+':'+str(l[1])) if l else '?':26s} instr {a[0]/div:8.1f} ({100*a[0]/T[0]:4.1f}%)  wavefronts {a[1]/div:7.1f} ({100*a[1]/max(T[1],1):4.1f}%) ideal {a[2]/div:7.1f}  stall {100*a[3]/max(T[3],1):4.1f}%")
