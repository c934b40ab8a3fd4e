"""The plant plug-in surface end to end: a plant header written by a user is compiled into its own library around
parallel-ddp_b200/csrc/plant_tu.cu, loaded with pddp_load_plant_library, and solved through the same C-ABI as the built-in plants."""
import os
import subprocess

import numpy as np
import pytest

from gpu_common import ROOT, pddp

pytestmark = pytest.mark.gpu
CSRC = os.path.join(ROOT, "parallel-ddp_b200", "csrc")


def _build(tmp_path, plant_id, header, incdir, name):
    so = str(tmp_path / f"libplant_{plant_id}.so")
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
                           f"-DPDDP_PLANT_ID={plant_id}", f"-DPDDP_PLANT_HEADER=\"{header}\"", f"-DPDDP_PLANT_NAME=\"{name}\"",
                           "-I", incdir, "-I", CSRC, os.path.join(CSRC, "plant_tu.cu"), "-o", so])
    return so


def test_external_plant_library_equals_builtin(tmp_path):
    """the pendulum header compiled as an EXTERNAL plug-in (plant number 11) gives bit for bit what the built-in PLANT 1 gives"""
    so = _build(tmp_path, 11, "plants/pendulum.cuh", CSRC, "pendulum as a plug-in library")
    assert pddp.load_plant_library(so) == 11 and pddp.plant_dims(11) == (1, 2, 1)
    x0, u0, xg = pddp.make_inputs(1, 32, 3, seed0=5)
    outs = []
    for plant in (1, 11):
        c = pddp.default_config(1, 32, 3, n_alpha=8, max_iter=25); c.plant = plant
        outs.append(pddp.Solver(c).runiLQR_GPU(x0, u0, xg))
    for k in ("x", "u", "Jout", "alphaOut", "iters"):
        assert np.array_equal(outs[0][k], outs[1][k], equal_nan=True), k


def test_user_written_plant(tmp_path):
    """a plant that exists nowhere in the library (pendulum with friction, model data through initI): the analytic integrator gradient
    agrees with finite differences of the integrator, the solver swings it up, and RK3 / Midpoint / Euler all run"""
    so = _build(tmp_path, 12, "damped_pendulum.cuh", os.path.join(ROOT, "tests", "plugin_example"), "damped pendulum")
    assert pddp.load_plant_library(so) == 12 and pddp.plant_dims(12) == (1, 2, 1)
    N, B = 64, 2
    x0 = np.zeros((B, N, 2), np.float32); x0[:, :, 1] = 0.001; u0 = np.full((B, N, 1), 0.01, np.float32)
    xg = np.tile(np.array([3.1416, 0.0], np.float32), (B, 1))
    for integ in (1, 2, 3):
        c = pddp.default_config(1, N, B, n_alpha=16, max_iter=60, integrator=integ); c.plant = 12
        s = pddp.Solver(c)
        # analytic AB against central differences of the integrator itself (float32: a loose tolerance)
        rng = np.random.default_rng(integ)
        xs = rng.normal(0, 1.0, (16, 2)).astype(np.float32); us = rng.normal(0, 2.0, (16, 1)).astype(np.float32)
        AB, qdd = s.integratorGradient(xs, us)
        assert np.allclose(qdd[:, 0], us[:, 0] - 9.81*np.sin(xs[:, 0]) - 0.3*xs[:, 1], rtol=1e-5, atol=1e-5)
        eps = 1e-2
        # (Euler only: the reference's Midpoint / RK3 gradients are not the exact derivatives of its own Midpoint / RK3 steps --
        # utils/integrators.cuh:78,181-191, reproduced as they are)
        for col in (range(3) if integ == 1 else ()):
            xp, xm, up, um = xs.copy(), xs.copy(), us.copy(), us.copy()
            if col < 2:
                xp[:, col] += eps; xm[:, col] -= eps
            else:
                up[:, 0] += eps; um[:, 0] -= eps
            fd = (s.integrator(xp, up).astype(np.float64) - s.integrator(xm, um)) / (2*eps)
            assert np.allclose(AB[:, col, :], fd, rtol=2e-2, atol=2e-3), (integ, col)
        out = s.runiLQR_GPU(x0, u0, xg)
        J = out["Jout"]
        assert np.isfinite(J[:, 0]).all() and (J[:, out["iters"][0]] < 0.2*J[:, 0]).all(), (integ, J[:, 0], J[:, -1])
        assert (np.abs(out["x"][:, -1, 0] - 3.1416) < 0.3).all(), out["x"][:, -1]
