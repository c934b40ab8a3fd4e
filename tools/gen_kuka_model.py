#!/usr/bin/env python3
"""Regenerate parallel-ddp_b200/csrc/kuka_model_data.inc from tests/golden/kuka_model.npz (robot model data)."""
import os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
z = np.load(os.path.join(ROOT, "tests", "golden", "kuka_model.npz"))


def fmt(v):
    s = "%.9g" % float(v)
    if "e" not in s and "." not in s:
        s += ".0"
    return s + "f"


out = [open(os.path.join(ROOT, "parallel-ddp_b200", "csrc", "kuka_model_data.inc")).read().split("static const")[0]]
for name, key in (("KUKA_I_DATA", "I"), ("KUKA_TBODY_DATA", "Tbody")):
    a = z[key]
    lines = ["    " + ", ".join(fmt(v) for v in a[r:r + 6]) + "," for r in range(0, len(a), 6)]
    out.append(f"static const float {name}[{len(a)}] = {{\n" + "\n".join(lines) + "\n};\n")
open(os.path.join(ROOT, "parallel-ddp_b200", "csrc", "kuka_model_data.inc"), "w").write("".join(out))
