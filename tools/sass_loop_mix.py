#!/usr/bin/env python3
"""Opcode-class mix of the innermost big loop of a kernel in libpddp.so (static count between the loop head label and its backward branch).
usage: tools/sass_loop_mix.py <mangled-substring>"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kern = sys.argv[1]
d = tempfile.mkdtemp(); subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "parallel-ddp_b200", "libpddp.so")], cwd=d, capture_output=True)
body = []
for f in os.listdir(d):
    s = subprocess.run(["nvdisasm", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
    cur = None
    for ln in s.splitlines():
        if ln.startswith(".text."): cur = ln; continue
        if cur and kern in cur: body.append(ln)
    if body: break
labels = {m.group(1): i for i, ln in enumerate(body) if (m := re.match(r"(\.L_x_\d+):", ln))}
best = None
for i, ln in enumerate(body):
    m = re.search(r"BRA\s+`\((\.L_x_\d+)\)", ln)
    if m and m.group(1) in labels and labels[m.group(1)] < i and (best is None or i - labels[m.group(1)] > best[1] - best[0]): best = (labels[m.group(1)], i)
CLS = {"FP": ("FFMA", "FMUL", "FADD", "MUFU", "FSEL", "FSETP", "FMNMX", "F2F", "F2I", "I2F", "FCHK"), "LDS": ("LDS",), "STS": ("STS",), "GMEM": ("LDG", "STG", "LDC", "LDCU", "LDL", "STL"),
       "SHFL": ("SHFL",), "SYNC": ("WARPSYNC", "BAR", "BSSY", "BSYNC", "NOP", "SYNCS"), "BRA": ("BRA", "EXIT", "CALL", "RET", "JMP", "BRX")}
tot = collections.Counter(); ops = collections.Counter()
for ln in body[best[0]:best[1]+1]:
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if not m: continue
    b = m.group(1).split(".")[0]; c = next((k for k, v in CLS.items() if b in v), "INT"); tot[c] += 1; ops[m.group(1)] += 1
print(sum(tot.values()), dict(tot)); print(ops.most_common(24))
