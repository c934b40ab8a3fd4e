"""The header shim: the reference example's call sequence (allocateMemory_GPU / runiLQR_GPU / freeMemory_GPU with the
reference's argument lists) compiled against libpddp.so must run and reproduce the Python binding's traces."""
import os
import re
import subprocess

import numpy as np
import pytest

from gpu_common import ROOT, pddp

pytestmark = pytest.mark.gpu


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
    o = s.runiLQR_GPU(x0, u0, xg)
    assert trace == list(o["alphaOut"][0])
    J0, J5 = (float(v) for v in re.search(r"J: ([\d.]+) -> ([\d.]+)", out).groups())
    assert abs(J0 - o["Jout"][0, 0]) < 1e-3 * J0 and abs(J5 - o["Jout"][0, 5]) < 1e-3 * J5
