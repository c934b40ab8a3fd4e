"""Pendulum / cart-pole / quadrotor (PLANT 1-3) and the Midpoint / RK3 integrators, CPU side.

The reference's plant files for these plants only compile through oracle/ref_harness/adapt_plant.cuh (they kept their v0.1
signatures); the fixtures tests/golden/p<plant>_i<integrator>_N<knots>_a<alphas>_{unit,trace}_H*.npz come from those builds of the
UNMODIFIED reference (tests/golden/make_goldens.py host).  Two things are pinned to them bit for bit, without a GPU:

* the oracle's host-arithmetic build (oracle/pddp_oracle_plants.c + the 1-D / 4-D Huu inverses and the rho retry of
  pddp_oracle.c): plant functions, integrator gradients, and every phase of whole 100-iteration solves;
* the plant headers the CUDA kernels are compiled from (parallel-ddp_b200/csrc/plants/*.cuh, plugin/integrators.cuh),
  instantiated for the host by tests/plant_host/plant_host.cu -- the same source the device code is generated from.
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
import trace_check

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
PLANT_HEADERS = {1: "pendulum", 2: "cartpole", 3: "quadrotor"}
UNIT_TAGS = ["p1_i3_N32_a1", "p1_i2_N32_a4", "p2_i3_N64_a8", "p2_i1_N32_a8", "p3_i3_N64_a16", "p3_i2_N32_a16"]
TRACES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "p?_i?_N*_trace_H_s*.npz")))
FP = C.POINTER(C.c_float)


def _gold(name):
    p = os.path.join(GOLD, name + ".npz")
    assert os.path.exists(p), f"golden fixture {name}.npz is missing (tests/golden/make_goldens.py host)"
    return dict(np.load(p))


def _weights(cfg):
    return np.array([cfg.Q1, cfg.Q2, cfg.R, cfg.QF1, cfg.QF2], np.float32)


@pytest.mark.parametrize("tag", UNIT_TAGS)
def test_oracle_plant_functions_vs_reference_host(tag):
    plant, integ, N, A = ol.parse_plant_tag(tag)
    d = _gold(tag + "_unit_H")
    L = ol.lib(False); cfg = ol.plant_cfg(plant, N, A, integ); cp = C.byref(cfg)
    n, m, npos = cfg.n, cfg.m, cfg.npos; nm = n + m; ns = int(d["meta"][3])
    assert np.float32(cfg.dt) == d["dt"][0]
    x = d["x"].reshape(ns, n); u = d["u"].reshape(ns, m); xg = d["xGoal"]
    for k in range(ns):
        q = np.zeros(npos, np.float32); L.orc_dynamics(cp, ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(q))
        assert np.array_equal(q, d["qdd"].reshape(ns, npos)[k]), (tag, k)
        AB = np.zeros(n*nm, np.float32); q2 = np.zeros(npos, np.float32)
        L.orc_integrator_gradient(cp, ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(AB), ol.fptr(q2))
        assert np.array_equal(AB, d["AB"].reshape(ns, -1)[k]), (tag, k)
        assert np.array_equal(q2, d["qdd_from_grad"].reshape(ns, npos)[k])
        assert np.float32(L.orc_cost(cp, ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(xg), 0)) == d["J_run"][k]
        assert np.float32(L.orc_cost(cp, ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(xg), N - 1)) == d["J_final"][k]
        H = np.zeros(nm*nm, np.float32); g = np.zeros(nm, np.float32)
        L.orc_cost_grad(cp, ol.fptr(H), ol.fptr(g), ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(xg), 0)
        assert np.array_equal(H, d["H_run"].reshape(ns, -1)[k]) and np.array_equal(g, d["g_run"].reshape(ns, -1)[k])
        L.orc_cost_grad(cp, ol.fptr(H), ol.fptr(g), ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(xg), N - 1)
        assert np.array_equal(H, d["H_final"].reshape(ns, -1)[k]) and np.array_equal(g, d["g_final"].reshape(ns, -1)[k])


@pytest.mark.parametrize("name", TRACES)
def test_oracle_whole_solve_vs_reference_host(name):
    """every phase of the dumped iterations + the complete Jout / alphaOut / x / u traces of 100 iterations, bit for bit"""
    tag = name.split("_trace_")[0]; plant, integ, N, A = ol.parse_plant_tag(tag)
    tr = _gold(name)
    tol = 0.0001 if name.endswith("_s2") else 0.0
    cfg = ol.plant_cfg(plant, N, A, integ, fma=False, tol_cost=tol, host_expred=True)
    assert np.array_equal(np.array(cfg.alpha[:A], np.float32), tr["alpha"])
    res, aOut, Jout = trace_check.run_trace_compare(tr, fma=False, host_expred=True, tol_cost=tol, cfg=cfg)
    bad = {k: v for k, v in res.items() if not v[0]}
    assert not bad, (name, list(bad.items())[:8])


def test_traces_cover_the_line_search_and_rejections():
    """the fixtures are not trivial: accepted steps with alpha index > 0 and rejected iterations both occur"""
    seen_pos, seen_rej = False, False
    for name in TRACES:
        a = _gold(name)["alphaOut"]
        seen_pos |= bool((a > 0).any()); seen_rej |= bool((a[1:] == -1).any())
    assert seen_pos and seen_rej


@pytest.fixture(scope="module")
def plant_host_libs(tmp_path_factory):
    out = {}
    d = tmp_path_factory.mktemp("plant_host")
    procs = []
    for pid, hdr in PLANT_HEADERS.items():
        so = os.path.join(d, f"libplant_host_{hdr}.so")
        procs.append((pid, so, subprocess.Popen(["nvcc", "-O3", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC", "-shared",
                                                  f"-DPDDP_PLANT_HEADER=\"plants/{hdr}.cuh\"", "-I", os.path.join(ROOT, "parallel-ddp_b200", "csrc"),
                                                  os.path.join(ROOT, "tests", "plant_host", "plant_host.cu"), "-o", so])))
    for pid, so, p in procs:
        assert p.wait() == 0
        L = C.CDLL(so); L.ph_cost.restype = C.c_float
        out[pid] = L
    return out


@pytest.mark.parametrize("tag", UNIT_TAGS)
def test_plugin_plant_headers_vs_reference_host(tag, plant_host_libs):
    """the plant headers and integrators the CUDA kernels are built from, compiled for the host, against the reference's host build"""
    plant, integ, N, A = ol.parse_plant_tag(tag)
    d = _gold(tag + "_unit_H"); L = plant_host_libs[plant]
    dims = (C.c_int*3)(); L.ph_dims(dims); npos, n, m = dims; nm = n + m
    cfg = ol.plant_cfg(plant, N, A, integ); w = _weights(cfg)
    assert (npos, n, m) == (cfg.npos, cfg.n, cfg.m)
    ns = int(d["meta"][3]); dt = C.c_float(d["dt"][0])
    x = d["x"].reshape(ns, n); u = d["u"].reshape(ns, m); xg = d["xGoal"]
    for k in range(ns):
        q = np.zeros(npos, np.float32); L.ph_dynamics(ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(q))
        assert np.array_equal(q, d["qdd"].reshape(ns, npos)[k]), (tag, k)
        AB = np.zeros(n*nm, np.float32); q2 = np.zeros(npos, np.float32)
        L.ph_integrator_gradient(integ, ol.fptr(x[k]), ol.fptr(u[k]), dt, ol.fptr(AB), ol.fptr(q2))
        assert np.array_equal(AB, d["AB"].reshape(ns, -1)[k]), (tag, k)
        assert np.array_equal(q2, d["qdd_from_grad"].reshape(ns, npos)[k])
        assert np.float32(L.ph_cost(N, ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(xg), 0, ol.fptr(w))) == d["J_run"][k]
        assert np.float32(L.ph_cost(N, ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(xg), N - 1, ol.fptr(w))) == d["J_final"][k]
        H = np.zeros(nm*nm, np.float32); g = np.zeros(nm, np.float32)
        L.ph_cost_grad(N, ol.fptr(H), ol.fptr(g), ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(xg), 0, ol.fptr(w))
        assert np.array_equal(H, d["H_run"].reshape(ns, -1)[k]) and np.array_equal(g, d["g_run"].reshape(ns, -1)[k])
        L.ph_cost_grad(N, ol.fptr(H), ol.fptr(g), ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(xg), N - 1, ol.fptr(w))
        assert np.array_equal(H, d["H_final"].reshape(ns, -1)[k]) and np.array_equal(g, d["g_final"].reshape(ns, -1)[k])
        # the integrator itself against the oracle's host build (the reference dumps no x_{k+1} in unit mode)
        xn = np.zeros(n, np.float32); xo = np.zeros(n, np.float32)
        L.ph_integrator(integ, ol.fptr(x[k]), ol.fptr(u[k]), dt, ol.fptr(xn))
        ol.lib(False).orc_integrator(C.byref(cfg), ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(xo))
        assert np.array_equal(xn, xo), (tag, k)


# ---------------------------------------------------------------------------------------------------------------- GPU arithmetic
# liboracle_fma.so against the reference's GPU run of these plants on a B200 (tests/golden/p*_{unit,trace,solve}_G*.npz; generated by
# tests/golden/make_goldens.py gpu on the GPU box, imported here): needs no device.
SOLVES_G = ["p1_i3_N32_a1", "p1_i2_N32_a4", "p2_i3_N64_a8", "p2_i1_N32_a8", "p3_i3_N64_a16", "p3_i2_N32_a16", "p3_i3_N256_a32"]
TRACES_G = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "p?_i?_N*_trace_G_s*.npz")))


@pytest.mark.parametrize("tag", UNIT_TAGS)
def test_oracle_gpu_arithmetic_plant_functions_vs_reference_gpu(tag):
    plant, integ, N, A = ol.parse_plant_tag(tag)
    g = _gold(tag + "_unit_G")
    L = ol.lib(True); cfg = ol.plant_cfg(plant, N, A, integ, fma=True); cp = C.byref(cfg)
    n, m, npos = cfg.n, cfg.m, cfg.npos; nm = n + m; ns = int(g["meta"][3])
    x = g["x"].reshape(ns, n); u = g["u"].reshape(ns, m)
    for k in range(ns):
        q = np.zeros(npos, np.float32); L.orc_dynamics(cp, ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(q))
        assert np.array_equal(q, g["qdd"].reshape(ns, npos)[k]), (tag, k)
        AB = np.zeros(n*nm, np.float32); q2 = np.zeros(npos, np.float32)
        L.orc_integrator_gradient(cp, ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(AB), ol.fptr(q2))
        assert np.array_equal(AB, g["AB"].reshape(ns, -1)[k]), (tag, k)
        assert np.array_equal(q2, g["qdd_from_grad"].reshape(ns, npos)[k]), (tag, k)


@pytest.mark.parametrize("name", TRACES_G)
def test_oracle_gpu_arithmetic_whole_solve_phases_vs_reference_gpu(name):
    tag = name.split("_trace_")[0]; plant, integ, N, A = ol.parse_plant_tag(tag)
    tr = _gold(name)
    tol = 0.0001 if name.endswith("_s2") else 0.0
    cfg = ol.plant_cfg(plant, N, A, integ, fma=True, tol_cost=tol, host_expred=False)
    res, aOut, Jout = trace_check.run_trace_compare(tr, fma=True, host_expred=False, tol_cost=tol, cfg=cfg, trig_limit=1.0e5)
    bad = {k: v for k, v in res.items() if not v[0]}
    assert not bad, (name, list(bad.items())[:8])


@pytest.mark.parametrize("tag", SOLVES_G)
def test_oracle_gpu_arithmetic_solves_vs_reference_gpu(tag):
    """100-iteration solves of the reference's GPU build, seeds 0-3: cost trace, step sizes, final trajectories"""
    import importlib
    pddp = importlib.import_module("parallel-ddp_b200")
    plant, integ, N, A = ol.parse_plant_tag(tag)
    g = _gold(tag + "_solve_G_s0-3")
    cfg = ol.plant_cfg(plant, N, A, integ, fma=True); L = ol.lib(True)
    x0, u0, xg = pddp.make_inputs(plant, N, 4, seed0=0)
    L1 = cfg.max_iter + 1; n, m = cfg.n, cfg.m
    for b in ((1, 2, 3) if N < 256 else (2,)):          # seed 0 draws the same numbers as seed 1 (minstd_rand0 maps 0 to 1)
        ox = np.zeros((N, n), np.float32); ou = np.zeros((N, m), np.float32)
        oJ = np.full(L1, np.nan, np.float32); oa = np.full(L1, -99, np.int32)
        L.orc_solve(C.byref(cfg), ol.fptr(x0[b]), ol.fptr(u0[b]), ol.fptr(xg[b]), ol.fptr(ox), ol.fptr(ou), ol.fptr(oJ), ol.iptr(oa))
        assert np.array_equal(oa, g["alphaOut"].reshape(4, L1)[b]), (tag, b, oa[:12], g["alphaOut"].reshape(4, L1)[b][:12])
        assert np.array_equal(oJ, g["Jout"].reshape(4, L1)[b], equal_nan=True), (tag, b)
        assert np.array_equal(ox, g["x_out"].reshape(4, N, n)[b]) and np.array_equal(ou, g["u_out"].reshape(4, N, m)[b]), (tag, b)
