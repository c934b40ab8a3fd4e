"""GPU parity of the plug-in plants: pendulum / cart-pole / quadrotor (PLANT 1-3) with Euler, Midpoint and RK3 (run with -m gpu).

Everything goes through the C-ABI.  Checkers, all bit for bit:
  (1) the reference's own GPU run of these plants on a B200 (fixtures p<plant>_i<integ>_N<knots>_a<alphas>_{unit,trace,solve}_G*.npz from
      oracle/_ref/ref_p*: the unmodified reference solver around its own plant files, oracle/ref_harness/adapt_plant.cuh);
  (2) the CPU oracle in GPU arithmetic (liboracle_fma.so), which the CPU suite pins to the reference's host build and
      test_oracle_* below pin to the reference's GPU run.
BASELINE.json configs covered: [0] pendulum N=32 alpha=1, [1] cart-pole N=64 alpha=8 batch 1, [4] quadrotor N=256 alpha=32 in the
receding-horizon loop (batch scaled to what the oracle finishes in seconds)."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import oracle_lib as ol
import trace_check
from gpu_common import ROOT, golden, pddp, report

pytestmark = pytest.mark.gpu
TAGS = ["p1_i3_N32_a1", "p1_i2_N32_a4", "p2_i3_N64_a8", "p2_i1_N32_a8", "p3_i3_N64_a16", "p3_i2_N32_a16"]
SOLVE_TAGS = TAGS + ["p3_i3_N256_a32"]


def _cfgs(tag, batch=1, fma=True, **over):
    plant, integ, N, A = ol.parse_plant_tag(tag)
    ocfg = ol.plant_cfg(plant, N, A, integ, fma=fma, tol_cost=over.get("tol_cost", 0.0))
    if "max_iter" in over:
        ocfg.max_iter = over["max_iter"]
    c = pddp.default_config(plant, N, batch, n_alpha=A, integrator=integ, **over)
    # the two default tables must agree (config.cuh per plant)
    for k in ("rho_init", "max_defect", "Q1", "Q2", "R", "QF1", "QF2"):
        assert np.float32(getattr(c, k)) == np.float32(getattr(ocfg, k)), k
    assert np.float32(ocfg.alpha[A-1]) == np.float32(float(c.alpha_base)**(A-1))
    return c, ocfg


def _oracle_solve(ocfg, x0, u0, xg):
    L = ol.lib(True); n, m, N, it = ocfg.n, ocfg.m, ocfg.N, ocfg.max_iter
    ox = np.zeros((N, n), np.float32); ou = np.zeros((N, m), np.float32)
    oJ = np.full(it + 1, np.nan, np.float32); oa = np.full(it + 1, -99, np.int32)
    iters = L.orc_solve(C.byref(ocfg), ol.fptr(x0), ol.fptr(u0), ol.fptr(xg), ol.fptr(ox), ol.fptr(ou), ol.fptr(oJ), ol.iptr(oa))
    return iters, ox, ou, oJ, oa


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


# ---------------------------------------------------------------------------------------------------------------- plant functions
@pytest.mark.parametrize("tag", TAGS)
def test_plant_functions_vs_reference_gpu_and_oracle(tag):
    """dynamics, _integratorGradient (AB), _integrator, costFunc, costGrad on the reference's unit samples"""
    plant, integ, N, A = ol.parse_plant_tag(tag)
    d = golden(tag + "_unit_H")                                          # inputs (the G dump uses the same seeded samples)
    g = golden(tag + "_unit_G")
    c, ocfg = _cfgs(tag)
    s = pddp.Solver(c); n, m, npos = s.n, s.m, s.npos; nm = n + m; ns = int(d["meta"][3])
    x = d["x"].reshape(ns, n); u = d["u"].reshape(ns, m); xg = d["xGoal"]
    assert _same(x, g["x"].reshape(ns, n)) and _same(u, g["u"].reshape(ns, m))
    qdd = s.dynamics(x, u); AB, qdd2 = s.integratorGradient(x, u); xn = s.integrator(x, u)
    J0, H0, g0 = s.cost(x, u, xg, 0); JN, HN, gN = s.cost(x, u, xg, N - 1)
    # (1) the reference's GPU run
    ok = dict(qdd=_same(qdd, g["qdd"].reshape(ns, npos)), AB=_same(AB.reshape(ns, -1), g["AB"].reshape(ns, -1)),
              qdd2=_same(qdd2, g["qdd_from_grad"].reshape(ns, npos)))
    # (2) the oracle in GPU arithmetic
    L = ol.lib(True); cp = C.byref(ocfg)
    oq = np.zeros((ns, npos), np.float32); oAB = np.zeros((ns, nm*n), np.float32); oxn = np.zeros((ns, n), np.float32)
    oJ0 = np.zeros(ns, np.float32); oJN = np.zeros(ns, np.float32); oH = np.zeros((ns, nm*nm), np.float32); og = np.zeros((ns, nm), np.float32)
    oHN = np.zeros((ns, nm*nm), np.float32); ogN = np.zeros((ns, nm), np.float32); q2 = np.zeros(npos, np.float32)
    for k in range(ns):
        L.orc_dynamics(cp, ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(oq[k]))
        L.orc_integrator_gradient(cp, ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(oAB[k]), ol.fptr(q2))
        L.orc_integrator(cp, ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(oxn[k]))
        oJ0[k] = L.orc_cost(cp, ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(xg), 0); oJN[k] = L.orc_cost(cp, ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(xg), N - 1)
        L.orc_cost_grad(cp, ol.fptr(oH[k]), ol.fptr(og[k]), ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(xg), 0)
        L.orc_cost_grad(cp, ol.fptr(oHN[k]), ol.fptr(ogN[k]), ol.fptr(x[k]), ol.fptr(u[k]), ol.fptr(xg), N - 1)
    ok.update(o_qdd=_same(qdd, oq), o_AB=_same(AB.reshape(ns, -1), oAB), o_xn=_same(xn, oxn), o_J=_same(J0, oJ0) and _same(JN, oJN),
              o_H=_same(H0.reshape(ns, -1), oH) and _same(HN.reshape(ns, -1), oHN), o_g=_same(g0, og) and _same(gN, ogN))
    nbad = dict(qdd=int(np.sum(qdd != oq)), AB=int(np.sum(AB.reshape(ns, -1) != oAB)), xn=int(np.sum(xn != oxn)), J=int(np.sum(J0 != oJ0) + np.sum(JN != oJN)))
    report(test="plant_functions", tag=tag, **ok, nbad_vs_oracle=nbad)
    assert all(ok.values()), (tag, ok, nbad)


# ---------------------------------------------------------------------------------------------------------------- whole solves
@pytest.mark.parametrize("tag", SOLVE_TAGS)
def test_solve_vs_reference_gpu(tag):
    """100-iteration solves of the reference's GPU build (seeds 0-3, TOL_COST 0), as ONE batch: cost trace, step-size trace,
    iteration counters and final trajectories bit for bit"""
    plant, integ, N, A = ol.parse_plant_tag(tag)
    g = golden(tag + "_solve_G_s0-3")
    B = 4
    c, _ = _cfgs(tag, batch=B)
    s = pddp.Solver(c)
    x0, u0, xg = pddp.make_inputs(plant, N, B, seed0=0)
    out = s.runiLQR_GPU(x0, u0, xg)
    L1 = c.max_iter + 1
    ok = dict(alphaOut=_same(out["alphaOut"], g["alphaOut"].reshape(B, L1)), Jout=_same(out["Jout"], g["Jout"].reshape(B, L1)),
              iters=_same(out["iters"], g["iters"].reshape(B)), x=_same(out["x"], g["x_out"].reshape(B, N, s.n)), u=_same(out["u"], g["u_out"].reshape(B, N, s.m)))
    report(test="plant_solve_vs_refG", tag=tag, **ok, J_end=[float(v) for v in out["Jout"][:, -1]])
    assert all(ok.values()), (tag, ok)


@pytest.mark.parametrize("tag,batch,iters,tol", [("p1_i3_N32_a1", 3, 40, 0.0), ("p1_i1_N64_a8", 2, 30, 0.0), ("p2_i3_N64_a8", 1, 60, 0.0),
                                                 ("p2_i2_N128_a16", 3, 25, 0.0001), ("p3_i3_N64_a16", 3, 30, 0.0), ("p3_i1_N32_a4", 2, 30, 0.0001),
                                                 ("p3_i3_N256_a32", 2, 12, 0.0)])
def test_solve_vs_oracle(tag, batch, iters, tol):
    """other shapes than the fixtures (other N, alpha counts, integrators, TOL_COST exits) against the oracle in GPU arithmetic"""
    plant, integ, N, A = ol.parse_plant_tag(tag)
    c, ocfg = _cfgs(tag, batch=batch, max_iter=iters, tol_cost=tol)
    s = pddp.Solver(c)
    x0, u0, xg = pddp.make_inputs(plant, N, batch, seed0=11)
    out = s.runiLQR_GPU(x0, u0, xg)
    for b in range(batch):
        it, ox, ou, oJ, oa = _oracle_solve(ocfg, x0[b], u0[b], xg[b])
        assert it == out["iters"][b], (tag, b, it, out["iters"][b])
        assert _same(oa, out["alphaOut"][b]), (tag, b, oa, out["alphaOut"][b])
        assert _same(oJ, out["Jout"][b]), (tag, b)
        assert _same(ox, out["x"][b]) and _same(ou, out["u"][b]), (tag, b)


def test_rho_retry_path_vs_oracle():
    """backwardPassGPU's retry (bpHelpers.cuh:497-511): a non-positive Huu makes a block fail, rho goes up, all blocks run again.
    Forced with a negative control weight R on the cart-pole (1-D inverse) and the quadrotor (4-D adjugate)."""
    for tag, R in (("p2_i3_N32_a8", -0.5), ("p3_i3_N32_a8", -3.0)):
        plant, integ, N, A = ol.parse_plant_tag(tag)
        c, ocfg = _cfgs(tag, batch=2, max_iter=12)
        c.R = R; ocfg.R = R
        s = pddp.Solver(c)
        x0, u0, xg = pddp.make_inputs(plant, N, 2, seed0=3)
        s.load_init(x0, u0, xg)
        rho0 = s.get("rho").copy()
        s.backwardPassGPU()
        assert (s.get("rho") > rho0).all(), "the scenario must exercise the retry"
        out = pddp.Solver(c).runiLQR_GPU(x0, u0, xg)
        for b in range(2):
            it, ox, ou, oJ, oa = _oracle_solve(ocfg, x0[b], u0[b], xg[b])
            assert it == out["iters"][b] and _same(oa, out["alphaOut"][b]) and _same(oJ, out["Jout"][b]), (tag, b, oa, out["alphaOut"][b])
            assert _same(ox, out["x"][b]) and _same(ou, out["u"][b])


def test_quadrotor_receding_horizon_vs_oracle():
    """BASELINE config[4] in shape (quadrotor, N=256, alpha=32, RK3, receding-horizon loop), batch and step count scaled to the oracle"""
    tag, B, steps, cap = "p3_i3_N256_a32", 2, 3, 4
    plant, integ, N, A = ol.parse_plant_tag(tag)
    c, ocfg = _cfgs(tag, batch=B, max_iter=cap, tol_cost=0.0001)
    s = pddp.Solver(c); n, m = s.n, s.m
    x0, u0, xg = pddp.make_inputs(plant, N, B, seed0=21)
    s.mpc_init(x0, u0)
    L = ol.lib(True); cp = C.byref(ocfg)
    mp = [L.orc_mpc_alloc(cp, ol.fptr(x0[b]), ol.fptr(u0[b]), ol.fptr(xg[b])) for b in range(B)]
    rng = np.random.default_rng(4)
    for step in range(steps):
        shift = np.array([1 + (step + b) % 3 for b in range(B)], np.int32)
        xa = np.stack([s.mpc_x[b, shift[b]] for b in range(B)]).astype(np.float32) + rng.normal(0, 0.01, (B, n)).astype(np.float32)
        out = s.mpc_step(xa, xg, shift, cap)
        for b in range(B):
            oJ = np.full(cap + 1, np.nan, np.float32); oa = np.full(cap + 1, -99, np.int32)
            it = L.orc_mpc_step(cp, mp[b], ol.fptr(xa[b]), ol.fptr(xg[b]), int(shift[b]), cap, 0, 0, ol.fptr(oJ), ol.iptr(oa))
            assert it == out["iters"][b], (step, b, it, out["iters"][b])
            assert _same(oa, out["alphaOut"][b][:cap+1]) and _same(oJ, out["Jout"][b][:cap+1]), (step, b, oa, out["alphaOut"][b])
            ox = np.ctypeslib.as_array(L.orc_mpc_x(mp[b]), shape=(N, n)); ou = np.ctypeslib.as_array(L.orc_mpc_u(mp[b]), shape=(N, m))
            oK = np.ctypeslib.as_array(L.orc_mpc_KT(mp[b]), shape=(N, n*m))
            assert _same(ox, out["x"][b]) and _same(ou, out["u"][b]) and _same(oK, out["KT"][b]), (step, b)
            assert L.orc_mpc_last_successful_solve(mp[b]) == out["last_successful_solve"][b]
    for q in mp:
        L.orc_mpc_free(q)


# ---------------------------------------------------------------------------------------------------------------- phase by phase
TRACES_G = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(ROOT, "tests", "golden", "p?_i?_N*_trace_G_s*.npz")))


@pytest.mark.parametrize("name", TRACES_G)
def test_phases_vs_reference_gpu_trace(name):
    """every array the reference's GPU run dumped in its first iterations (P, p, KT, du, A-BK, Bdu, expected reduction, swept and simulated
    x, u, d of every step size, J, defects, AB, H, g, rho schedule), phase by phase through the phase-level entry points"""
    from test_gpu_parity import _phase_walk
    tag = name.split("_trace_")[0]; plant, integ, N, A = ol.parse_plant_tag(tag)
    tol = 0.0001 if name.endswith("_s2") else 0.0
    c, _ = _cfgs(tag, batch=1, tol_cost=tol)
    res = _phase_walk(golden(name), tol, s=pddp.Solver(c))
    bad = [k for k, v in res.items() if not v[0]]
    report(test="plant_phases_vs_refG", golden=name, nchecks=len(res), inexact=bad[:20])
    assert not bad, bad[:10]
