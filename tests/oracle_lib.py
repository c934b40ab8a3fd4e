"""ctypes binding of the CPU oracle (oracle/liboracle*.so).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg -- never by the product package."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
MAX_ALPHA = 64
F = C.c_float
FP = C.POINTER(C.c_float)
IP = C.POINTER(C.c_int)


EE_WEIGHT_NAMES = ("Q_EE1", "Q_EE2", "QF_EE1", "QF_EE2", "R_EE", "Q_xdEE", "QF_xdEE", "Q_xEE", "QF_xEE")


class Cfg(C.Structure):
    _fields_ = [("plant", C.c_int), ("n", C.c_int), ("m", C.c_int), ("npos", C.c_int), ("N", C.c_int),
                ("n_alpha", C.c_int), ("M", C.c_int), ("integrator", C.c_int), ("max_iter", C.c_int),
                ("expred_host_order", C.c_int), ("dt", F), ("alpha", F * MAX_ALPHA),
                ("rho_init", F), ("rho_min", F), ("rho_max", F), ("rho_factor", F),
                ("exp_red_min", F), ("exp_red_max", F), ("max_defect", F), ("tol_cost", F),
                ("Q1", F), ("Q2", F), ("R", F), ("QF1", F), ("QF2", F),
                ("I", F * 252), ("Tbody", F * 252), ("gravity", C.c_float), ("ee_cost", C.c_int)] + \
               [(k, F) for k in ("Q_EE1", "Q_EE2", "QF_EE1", "QF_EE2", "R_EE", "Q_xdEE", "QF_xdEE", "Q_xEE", "QF_xEE")] + \
               [("use_xtarget", C.c_int), ("xTarget", F * 16), ("final_cost_shift", C.c_int),
                ("use_limits", C.c_int), ("Q_PL", F), ("Q_VL", F), ("R_TL", F),
                ("use_smooth_abs", C.c_int), ("sa_alpha", F), ("sa_alpha2", F)]


class Ws(C.Structure):
    _fields_ = [(k, FP) for k in ("x", "u", "d", "xp", "xp2", "up", "dp", "AB", "H", "g", "P", "p", "Pp", "pp",
                                   "KT", "du", "ApBK", "Bdu", "xg")] + \
               [("J", F * MAX_ALPHA), ("dT", F * MAX_ALPHA), ("dJexp", F * 32), ("err", C.c_int * 16),
                ("prevJ", F), ("dJ", F), ("z", F), ("rho", F), ("drho", F),
                ("iter", C.c_int), ("alphaIndex", C.c_int), ("ignore_defect", C.c_int), ("JTp", F * (MAX_ALPHA * 16))]


def _build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def fptr(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(FP)


def iptr(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(IP)


_libs = {}


def lib(fma=False):
    name = "liboracle_fma.so" if fma else "liboracle.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(ORACLE_DIR, name)
    src = os.path.join(ORACLE_DIR, "pddp_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        _build()
    L = C.CDLL(path)
    cp = C.POINTER(Cfg)
    wp = C.POINTER(Ws)
    L.orc_default_cfg_kuka.argtypes = [cp, C.c_int]
    L.orc_default_cfg_plant.argtypes = [cp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_dynamics.argtypes = [cp, FP, FP, FP]
    L.orc_dynamics_gradient_any.argtypes = [cp, FP, FP, FP, FP]
    L.orc_ws_alloc.argtypes = [cp]; L.orc_ws_alloc.restype = wp
    L.orc_ws_free.argtypes = [wp]
    L.orc_kuka_dynamics.argtypes = [cp, FP, FP, FP]
    L.orc_kuka_dynamics_gradient.argtypes = [cp, FP, FP, FP, FP]
    L.orc_integrator.argtypes = [cp, FP, FP, FP]
    L.orc_integrator_gradient.argtypes = [cp, FP, FP, FP, FP]
    L.orc_cost.argtypes = [cp, FP, FP, FP, C.c_int]; L.orc_cost.restype = F
    L.orc_cost_grad.argtypes = [cp, FP, FP, FP, FP, FP, C.c_int]
    L.orc_load.argtypes = [cp, wp, FP, FP, FP]
    L.orc_init.argtypes = [cp, wp, FP, IP]
    L.orc_backward_pass.argtypes = [cp, wp]; L.orc_backward_pass.restype = C.c_int
    L.orc_backward_pass_once.argtypes = [cp, wp, F]
    for fn in ("orc_forward_sweep", "orc_forward_sim", "orc_cost_defect", "orc_line_search", "orc_next_iteration_setup"):
        getattr(L, fn).argtypes = [cp, wp]
    L.orc_accept_reject.argtypes = [cp, wp, FP, IP]; L.orc_accept_reject.restype = C.c_int
    L.orc_solve.argtypes = [cp, FP, FP, FP, FP, FP, FP, IP]; L.orc_solve.restype = C.c_int
    L.orc_mpc_alloc.argtypes = [cp, FP, FP, FP]; L.orc_mpc_alloc.restype = C.c_void_p
    L.orc_mpc_free.argtypes = [C.c_void_p]
    L.orc_mpc_step.argtypes = [cp, C.c_void_p, FP, FP, C.c_int, C.c_int, C.c_int, C.c_int, FP, IP]; L.orc_mpc_step.restype = C.c_int
    for _f in ('orc_mpc_x', 'orc_mpc_u', 'orc_mpc_KT'):
        getattr(L, _f).argtypes = [C.c_void_p]; getattr(L, _f).restype = FP
    L.orc_mpc_last_successful_solve.argtypes = [C.c_void_p]; L.orc_mpc_last_successful_solve.restype = C.c_int
    L.orc_solve_ex.argtypes = [cp, FP, FP, FP, FP, FP, FP, FP, C.c_int, C.c_int, C.c_int, FP, FP, FP, IP]; L.orc_solve_ex.restype = C.c_int
    L.orc_ee_pos.argtypes = [cp, FP, FP, FP]
    L.orc_ee_cost.argtypes = [cp, FP, FP, FP, FP, C.c_int]; L.orc_ee_cost.restype = F
    L.orc_ee_cost_grad.argtypes = [cp, FP, FP, FP, FP, FP, FP, FP, C.c_int]
    L.orc_fma_mode.restype = C.c_int
    assert L.orc_fma_mode() == (1 if fma else 0)
    _libs[name] = L
    return L


def kuka_model():
    """Kuka iiwa14 spatial inertias and fixed joint transforms (robot model data, tests/golden/kuka_model.npz,
    dumped from the reference's initI/initT by tests/golden/make_goldens.py)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "kuka_model.npz"))
    return z["I"].astype(np.float32), z["Tbody"].astype(np.float32)


def kuka_cfg(N, fma=False, tol_cost=0.0, host_expred=False, ee_weights=None, x_target=None):
    """ee_weights: the nine end-effector cost weights (Q_EE1, Q_EE2, QF_EE1, QF_EE2, R_EE, Q_xdEE, QF_xdEE, Q_xEE, QF_xEE) -> EE_COST 1"""
    L = lib(fma)
    c = Cfg()
    L.orc_default_cfg_kuka(C.byref(c), N)
    I, Tb = kuka_model()
    c.I[:] = list(I); c.Tbody[:] = list(Tb)
    c.tol_cost = tol_cost
    c.expred_host_order = 1 if host_expred else 0
    if ee_weights is not None:
        c.ee_cost = 1
        for k, v in zip(EE_WEIGHT_NAMES, ee_weights):
            setattr(c, k, float(v))
    if x_target is not None:
        c.use_xtarget = 1
        c.xTarget[:14] = [float(v) for v in x_target]
    return c


def plant_cfg(plant, N, n_alpha, integrator, fma=False, tol_cost=0.0, host_expred=False):
    """PLANT 1-3 (pendulum, cart-pole, quadrotor) with the reference's defaults (config.cuh:21-61 and the plants' cost weights)."""
    L = lib(fma)
    c = Cfg()
    L.orc_default_cfg_plant(C.byref(c), plant, N, n_alpha, integrator)
    c.tol_cost = tol_cost
    c.expred_host_order = 1 if host_expred else 0
    return c


def parse_plant_tag(tag):
    """'p2_i3_N64_a8' -> (plant, integrator, N, n_alpha): the naming of oracle/Makefile's PLANT 1-3 reference builds"""
    p, i, n, a = tag.split("_")[:4]
    return int(p[1:]), int(i[1:]), int(n[1:]), int(a[1:])


class WsView:
    """numpy views over an orc_ws allocated by the oracle."""

    def __init__(self, L, cfg):
        self.L, self.cfg = L, cfg
        self.ptr = L.orc_ws_alloc(C.byref(cfg))
        n, m, N, A = cfg.n, cfg.m, cfg.N, cfg.n_alpha
        nm = n + m
        shapes = dict(x=(A, N, n), u=(A, N, m), d=(A, N, n), xp=(N, n), xp2=(N, n), up=(N, m), dp=(N, n),
                      AB=(N, nm, n), H=(N, nm, nm), g=(N, nm), P=(N, n, n), p=(N, n), Pp=(N, n, n), pp=(N, n),
                      KT=(N, m, n), du=(N, m), ApBK=(N, n, n), Bdu=(N, n), xg=(n,))
        for k, shp in shapes.items():
            setattr(self, k, np.ctypeslib.as_array(getattr(self.ptr.contents, k), shape=shp))
        self.shapes = shapes

    @property
    def s(self):
        return self.ptr.contents

    def free(self):
        self.L.orc_ws_free(self.ptr)


# ---- double-precision build (finite-difference reference only) ----------------------------------
D = C.c_double
DP = C.POINTER(C.c_double)


class Cfg64(C.Structure):
    _fields_ = [(n, (D if t is F else (D * t._length_ if hasattr(t, "_length_") and t._type_ is F else t))) for n, t in Cfg._fields_]


def lib64():
    if "f64" in _libs:
        return _libs["f64"]
    path = os.path.join(ORACLE_DIR, "liboracle_f64.so")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(ORACLE_DIR, "pddp_oracle.c")):
        _build()
    L = C.CDLL(path)
    cp = C.POINTER(Cfg64)
    L.orc_default_cfg_kuka.argtypes = [cp, C.c_int]
    L.orc_integrator.argtypes = [cp, DP, DP, DP]
    L.orc_integrator_gradient.argtypes = [cp, DP, DP, DP, DP]
    _libs["f64"] = L
    return L


def kuka_cfg64(N):
    L = lib64()
    c = Cfg64()
    L.orc_default_cfg_kuka(C.byref(c), N)
    I, Tb = kuka_model()
    c.I[:] = [float(v) for v in I]; c.Tbody[:] = [float(v) for v in Tb]
    return c


def dptr(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(DP)
