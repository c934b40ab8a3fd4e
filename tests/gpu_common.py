import importlib
import json
import os

import numpy as np
import pytest

import refdump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pddp = importlib.import_module("parallel-ddp_b200")
REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")


def golden(name, required=True):
    """tests/golden/<name>.npz, or the raw dump written earlier in the same gpurun call (gpurun_out/golden_raw/<name>.bin)."""
    p = os.path.join(ROOT, "tests", "golden", name + ".npz")
    if os.path.exists(p):
        return dict(np.load(p))
    p = os.path.join(ROOT, "gpurun_out", "golden_raw", name + ".bin")
    if os.path.exists(p):
        return refdump.load(p)
    if required:
        pytest.fail(f"golden fixture {name} is missing: tests/golden/{name}.npz (tests/golden/make_goldens.py) -- a lost fixture must not pass silently")
    pytest.skip(f"golden {name} not generated yet")


def relerr(mine, ref):
    mine = np.asarray(mine, np.float64); ref = np.asarray(ref, np.float64).reshape(mine.shape)
    return float(np.max(np.abs(mine - ref)) / (np.max(np.abs(ref)) + 1e-30)) if mine.size else 0.0


def report(**kw):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(json.dumps(kw) + "\n")
