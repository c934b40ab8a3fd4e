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
    for kw in (dict(plant=1), dict(N=100), dict(M=3), dict(n_alpha=0), dict(integrator=3)):
        c = pddp.default_config_kuka(32, 1)
        for k, v in kw.items():
            setattr(c, k, v)
        h = C.c_void_p()
        assert lib.pddp_create(C.byref(c), C.byref(h)) == -1
        assert lib.pddp_last_error(None)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(pddp.PddpError, match="no CUDA device"):
        pddp.Solver(pddp.default_config_kuka(32, 1))
