#!/usr/bin/env python3
"""Static instruction count per source line of one kernel (nvdisasm -g -c of the cubin inside libpddp.so), by opcode class.
usage: tools/sass_lines.py <kernel-substring> [cubin-name-substring]  -- e.g. sim_kernelILb0ELi16 pddp_api"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "parallel-ddp_b200", "libpddp.so")
kern = sys.argv[1]; cub = sys.argv[2] if len(sys.argv) > 2 else "pddp_api"
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
f = [x for x in os.listdir(d) if cub in x][0]
out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
CLS = {"FP": ("FFMA", "FMUL", "FADD", "MUFU", "FSEL", "FSETP", "FMNMX", "DFMA", "DMUL", "DADD", "F2F", "F2I", "I2F", "FCHK"), "LDS": ("LDS",), "STS": ("STS",), "GMEM": ("LDG", "STG", "LDC", "LDCU", "LDL", "STL"),
       "SHFL": ("SHFL",), "SYNC": ("WARPSYNC", "BAR", "BSSY", "BSYNC", "NOP", "SYNCS"), "BRA": ("BRA", "EXIT", "CALL", "RET", "JMP")}
def cls(op):
    b = op.split(".")[0]
    for k, v in CLS.items():
        if b in v: return k
    return "INT"
cur = None; line = None; per = collections.defaultdict(collections.Counter); tot = collections.Counter()
for ln in out.splitlines():
    if ln.startswith(".text."):
        cur = ln; continue
    if cur is None or kern not in cur: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m: c = cls(m.group(1)); per[line][c] += 1; tot[c] += 1
keys = ["FP", "LDS", "STS", "GMEM", "SHFL", "INT", "SYNC", "BRA"]
print(f"{'line':28s} {'total':>6s} " + " ".join(f"{k:>6s}" for k in keys))
for l, h in sorted(per.items(), key=lambda kv: (kv[0] or ("", 0))):
    print(f"{(l[0] + ':' + str(l[1])) if l else '?':28s} {sum(h.values()):6d} " + " ".join(f"{h.get(k, 0):6d}" for k in keys))
print(f"{'ALL':28s} {sum(tot.values()):6d} " + " ".join(f"{tot.get(k, 0):6d}" for k in keys))
