"""CPU-side checks of the drop-in boundary: libpddp.so loads, exports every symbol include/pddp.h declares, and refuses to
run without a device (no CPU fallback).  No compute is launched here."""
import ctypes as C
import importlib
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pddp = importlib.import_module("parallel-ddp_b200")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(pddp.LIB_PATH):
        subprocess.check_call([os.path.join(ROOT, "parallel-ddp_b200", "build.sh")])
    return pddp.load_library()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "pddp.h")).read()
    names = sorted(set(re.findall(r"\b(pddp_[a-z_]+)\s*\(", hdr)))
    assert len(names) >= 19
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(names) == set(pddp.EXPORTS)


def test_config_struct_matches_header(lib):
    c = pddp.default_config_kuka(128, 64)
    assert (c.plant, c.N, c.n_alpha, c.M, c.max_iter, c.batch, c.integrator) == (4, 128, 16, 4, 100, 64, 1)
    assert abs(c.rho_init - 12.5) < 1e-6 and abs(c.QF2 - 1000.0) < 1e-3 and abs(c.exp_red_max - 1.25) < 1e-6


def test_inputs_match_reference_harness(lib, golden_dir):
    """pddp_make_inputs_kuka reproduces the deterministic harness' x0/u0/goal (same libstdc++ engine and distribution)."""
    import numpy as np
    tr = dict(np.load(os.path.join(golden_dir, "trace_H_N32_s3_tol1e-4.npz")))
    x0, u0, xg = pddp.make_inputs_kuka(32, 5, seed0=0)
    assert np.array_equal(x0[3].ravel(), tr["x_in"]) and np.array_equal(u0[3].ravel(), tr["u_in"]) and np.array_equal(xg[3], tr["xGoal"])


def test_invalid_configs_are_rejected(lib):
    for kw in (dict(plant=9), dict(N=100), dict(M=3), dict(n_alpha=0), dict(integrator=3)):
        c = pddp.default_config_kuka(32, 1)
        for k, v in kw.items():
            setattr(c, k, v)
        h = C.c_void_p()
        assert lib.pddp_create(C.byref(c), C.byref(h)) == -1
        assert lib.pddp_last_error(None)
    for plant, kw in ((1, dict(integrator=4)), (2, dict(ee_cost=1)), (3, dict(integrator=0))):
        c = pddp.default_config(plant, 32, 1, **kw)
        h = C.c_void_p()
        assert lib.pddp_create(C.byref(c), C.byref(h)) == -1


def test_plant_tables_and_defaults(lib):
    """the built-in plug-in plants are registered with the reference's dimensions and defaults (config.cuh:21-61)"""
    assert pddp.plant_dims(1) == (1, 2, 1) and pddp.plant_dims(2) == (2, 4, 1) and pddp.plant_dims(3) == (6, 12, 4) and pddp.plant_dims(4) == (7, 14, 7)
    with pytest.raises(pddp.PddpError):
        pddp.plant_dims(17)
    c = pddp.default_config(2, 64, 1)
    assert (c.integrator, c.n_alpha, c.M) == (3, 32, 4) and abs(c.max_defect - 0.75) < 1e-7 and abs(c.rho_init - 10.0) < 1e-6 and abs(c.alpha_base - 0.75) < 1e-7
    c = pddp.default_config(3, 256, 1)
    assert (c.n_alpha, c.integrator) == (16, 3) and abs(c.R - 5.0) < 1e-6 and abs(c.total_time - 4.0) < 1e-6
    assert lib.pddp_load_plant_library(b"/nonexistent/libplant.so") < 0 and b"dlopen" in lib.pddp_plant_error()


def test_plant_inputs_match_reference_harness(lib, golden_dir):
    """pddp_make_inputs for PLANT 1-3 = the deterministic harness' x0 / u0 / goal (WAFR_iLQR_examples.cu:19-33,72-78,87-90,110-115)"""
    import numpy as np
    for plant, name, seed in ((1, "p1_i3_N32_a1_trace_H_s0", 0), (2, "p2_i3_N64_a8_trace_H_s2", 2), (3, "p3_i2_N32_a16_trace_H_s1", 1)):
        tr = dict(np.load(os.path.join(golden_dir, name + ".npz")))
        N = int(tr["meta"][0])
        x0, u0, xg = pddp.make_inputs(plant, N, 3, seed0=0)
        assert np.array_equal(x0[seed].ravel(), tr["x_in"]) and np.array_equal(u0[seed].ravel(), tr["u_in"]) and np.array_equal(xg[seed], tr["xGoal"])


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(pddp.PddpError, match="no CUDA device"):
        pddp.Solver(pddp.default_config_kuka(32, 1))
