"""The header shim (include/pddp_shim.cuh), SURVEY 8b.1.

* test_reference_example_*: the reference's own examples/WAFR_iLQR_examples.cu, unmodified except for its `#include "../config.cuh"`
  line, built against libpddp.so by `make -C oracle example_shim` (in the container that has /root/reference; the binaries travel to
  the GPU box in oracle/_ref/) and run here in mode G.  The example seeds with time(0), so format and finiteness are checked, not values.
* test_shim_example_*: the example's call sequence in a small self-contained program with fixed inputs, compared with the Python binding."""
import os
import re
import subprocess

import numpy as np
import pytest

from gpu_common import ROOT, pddp

pytestmark = pytest.mark.gpu
LINE = re.compile(r"GPU Parallel blocks:\[(\d+)\] t:\[([-\d.e+naninf]+)\] with FP\[([-\d.e+]+)\], FS\[([-\d.e+]+)\], BP\[([-\d.e+]+)\], NIU\[([-\d.e+]+)\] "
                  r"Xf:\[([-\d.e+naninf]+), ([-\d.e+naninf]+)\] iters:\[(\d+)\] cost:\[([-\d.e+naninf]+)\] max_d\[([-\d.e+naninf]+)\]")


@pytest.mark.parametrize("exe,iters_expected", [("example_shim_p4", 100), ("example_shim_p2", 100)])
def test_reference_example_runs_against_the_shim(exe, iters_expected):
    path = os.path.join(ROOT, "oracle", "_ref", exe)
    assert os.path.exists(path), f"{path} is missing: make -C oracle example_shim (needs /root/reference)"
    r = subprocess.run([path, "G"], input="q\n", capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [LINE.search(l) for l in r.stdout.splitlines() if l.startswith("GPU Parallel blocks")]
    assert len(lines) == 100 and all(lines), "one well-formed summary line per solve (TEST_ITERS 100, DDPWrappers.cuh:134)"
    for m in lines:
        assert int(m.group(1)) == 4 and int(m.group(9)) == iters_expected            # M_BLOCKS, iterations (TOL_COST 0: all 100)
        vals = [float(m.group(i)) for i in (2, 3, 4, 5, 6, 7, 8, 10, 11)]
        assert np.isfinite(vals).all() and vals[0] > 0 and all(v > 0 for v in vals[1:5]) and vals[7] > 0 and vals[8] >= 0
    # the example's own statistics: 101 cost rows, a median time trace with MAX_ITER + 1 increasing entries
    assert len(re.findall(r"^Iter \d+: Median\[", r.stdout, re.M)) == 101
    trace = [float(v) for v in r.stdout.split("Median Time Trace:\n")[1].splitlines()[0].split()]
    assert len(trace) == 101 and trace[0] == 0.0 and all(b > a for a, b in zip(trace, trace[1:])), "per-iteration timing arrays are filled for every iteration"
    costs = [float(v) for v in r.stdout.split("Median J Trace:\n")[1].splitlines()[0].split()]
    assert len(costs) == 101 and costs[-1] < costs[0]


def test_reference_example_cpu_modes_stop_with_a_message():
    path = os.path.join(ROOT, "oracle", "_ref", "example_shim_p4")
    assert os.path.exists(path)
    for mode in ("C", "CS", "S"):
        r = subprocess.run([path, mode], input="q\n", capture_output=True, text=True, timeout=60)
        assert r.returncode == 2 and "not built" in r.stderr


@pytest.mark.parametrize("ee", [0, 1])
def test_shim_example_runs_and_matches(tmp_path, ee):
    exe = str(tmp_path / "shim_example")
    libdir = os.path.join(ROOT, "parallel-ddp_b200")
    subprocess.check_call(["nvcc", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", f"-DEE_COST={ee}", "-o", exe, os.path.join(ROOT, "tests", "shim_example.cu"),
                           "-L", libdir, "-lpddp", "-Xlinker", "-rpath", "-Xlinker", libdir])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    assert "GPU Parallel blocks:[4]" in out and "iters:[5]" in out
    trace = [int(v) for v in re.search(r"alpha trace:([-\d ]+)", out).group(1).split()]
    x0, u0, xg = pddp.make_inputs_kuka(32, 1, 0)
    if ee:      # the example's goal pose, the reference's default weights (cost_arm.cuh:106-117)
        xg[:] = 0; xg[0, :6] = np.array([0.3638, 0.0, 1.0628, 0.5 * 3.14159, 0.0, 0.5 * 3.14159], np.float32)
    s = pddp.Solver(pddp.default_config_kuka(32, 1, max_iter=5, ee_cost=ee))
    o = s.runiLQR_GPU(x0, u0, xg, want_times=True)
    assert trace == list(o["alphaOut"][0])
    J0, J5 = (float(v) for v in re.search(r"J: ([\d.]+) -> ([\d.]+)", out).groups())
    assert abs(J0 - o["Jout"][0, 0]) < 1e-3 * J0 and abs(J5 - o["Jout"][0, 5]) < 1e-3 * J5
    # max_d of the summary line is the final trajectory's defect, and the per-iteration timing arrays are filled
    md = float(LINE.search(out).group(11))
    assert abs(md - float(s.final_max_defect()[0])) <= 1e-6 * max(1.0, md)
    t = s.iteration_times()
    assert all(len(t[k]) == 5 and (t[k] > 0).all() for k in ("sim", "sweep", "bp", "nis"))


def test_phase_times_need_one_group():
    """times_ms[1..4] and the per-iteration times exist for one problem group; with several groups the phases overlap and they are 0"""
    x0, u0, xg = pddp.make_inputs_kuka(32, 16, 0)
    s = pddp.Solver(pddp.default_config_kuka(32, 16, max_iter=4))
    assert s.set_groups(4) == 4
    o = s.runiLQR_GPU(x0, u0, xg, want_times=True)
    assert o["times_ms"]["total"] > 0 and o["times_ms"]["bp"] == 0
    assert s.set_groups(1) == 1
    o = s.runiLQR_GPU(x0, u0, xg, want_times=True)
    assert all(o["times_ms"][k] > 0 for k in ("total", "sim", "sweep", "bp", "nis", "init"))
    assert len(s.iteration_times()["bp"]) == 4
