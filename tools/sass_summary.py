#!/usr/bin/env python3
"""Opcode histogram per kernel of libpddp.so (cuobjdump -sass): python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "parallel-ddp_b200", "libpddp.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}   (arch: {', '.join(arch)})")
print("# per kernel: instruction count, then the opcodes that matter for the profiling recipe -- FP32 (FFMA/FMUL/FADD), FP64, MUFU, shared memory")
print("# (LDS/STS), global (LDG/STG), warp shuffles, 1-D bulk TMA (UBLKCP) and mbarrier (SYNCS), block barriers (BAR), tensor-core ops (none expected:")
print("# SURVEY 8d rules them out)")
cur = None; hist = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
KEYS = ["FFMA", "FMUL", "FADD", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "UBLKCP", "SYNCS", "BAR", "WARPSYNC", "IMAD", "LOP3", "ISETP", "BRA", "UTMALDG", "UTCHMMA", "HMMA"]
def demangle(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\(.*", "", r)[:70]
print(f"{'kernel':72s} {'total':>7s} " + " ".join(f"{k:>7s}" for k in KEYS))
for k, h in sorted(hist.items(), key=lambda kv: -sum(kv[1].values())):
    tot = sum(h.values())
    print(f"{demangle(k):72s} {tot:7d} " + " ".join(f"{h.get(x, 0):7d}" for x in KEYS))
tot = collections.Counter()
for h in hist.values():
    tot.update(h)
print(f"{'ALL':72s} {sum(tot.values()):7d} " + " ".join(f"{tot.get(x, 0):7d}" for x in KEYS))
