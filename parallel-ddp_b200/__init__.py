"""parallel-ddp_b200 -- host-side mirror of the reference's solver surface over libpddp.so (C-ABI, include/pddp.h).

The reference is header-only C++ (DDPHelpers/DDPWrappers.cuh); a Python binding does not exist upstream, so this module
mirrors its entry points by name and argument meaning:

    allocateMemory_GPU(...)  -> Solver(cfg)            nisInitHelpers.cuh:766
    runiLQR_GPU(...)         -> Solver.runiLQR_GPU()   DDPWrappers.cuh:8
    freeMemory_GPU(...)      -> Solver.freeMemory_GPU() nisInitHelpers.cuh:863

There is NO CPU fallback: importing works anywhere (so the symbol table can be checked without a GPU), but creating a
Solver without the compiled CUDA library or without a CUDA device raises PddpError.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpddp.so")
MAX_ALPHA = 32

PLANT_PEND, PLANT_CART, PLANT_QUAD, PLANT_KUKA = 1, 2, 3, 4


class PddpError(RuntimeError):
    pass


# the nine weights of the end-effector cost, in the order of pddp_config (plants/cost_arm.cuh:106-115)
EE_WEIGHT_NAMES = ("Q_EE1", "Q_EE2", "QF_EE1", "QF_EE2", "R_EE", "Q_xdEE", "QF_xdEE", "Q_xEE", "QF_xEE")


class Config(C.Structure):
    """pddp_config (include/pddp.h): run-time form of config.cuh."""
    _fields_ = [("plant", C.c_int), ("N", C.c_int), ("n_alpha", C.c_int), ("M", C.c_int), ("max_iter", C.c_int),
                ("batch", C.c_int), ("device", C.c_int), ("integrator", C.c_int),
                ("alpha_base", C.c_float), ("total_time", C.c_float),
                ("rho_init", C.c_float), ("rho_min", C.c_float), ("rho_max", C.c_float), ("rho_factor", C.c_float),
                ("exp_red_min", C.c_float), ("exp_red_max", C.c_float), ("max_defect", C.c_float), ("tol_cost", C.c_float),
                ("Q1", C.c_float), ("Q2", C.c_float), ("R", C.c_float), ("QF1", C.c_float), ("QF2", C.c_float), ("gravity", C.c_float),
                ("ee_cost", C.c_int)] + [(k, C.c_float) for k in EE_WEIGHT_NAMES] + \
               [("use_limits", C.c_int), ("lim_Q_pos", C.c_float), ("lim_Q_vel", C.c_float), ("lim_R_tau", C.c_float),
                ("use_smooth_abs", C.c_int), ("smooth_abs_alpha", C.c_double)]


EXPORTS = ["pddp_default_config_kuka", "pddp_create", "pddp_destroy", "pddp_last_error", "pddp_solve", "pddp_solve_device",
           "pddp_make_inputs_kuka", "pddp_unit_dynamics", "pddp_unit_integrator_gradient", "pddp_set_array", "pddp_get_array",
           "pddp_phase_load_init", "pddp_phase_backward_pass", "pddp_phase_forward_sweep", "pddp_phase_forward_sim",
           "pddp_phase_line_search", "pddp_phase_next_iteration", "pddp_last_phase_stats", "pddp_last_launch_count", "pddp_set_groups", "pddp_selftest_rcp", "pddp_selftest_sincos", "pddp_set_warm_start", "pddp_set_start_mode", "pddp_mpc_init", "pddp_mpc_step", "pddp_set_skip_unchanged", "pddp_set_x_target", "pddp_mpc_set_cost_shift",
           "pddp_default_config", "pddp_plant_dims", "pddp_register_plant", "pddp_load_plant_library", "pddp_plant_error", "pddp_make_inputs",
           "pddp_unit_integrator", "pddp_unit_cost", "pddp_last_iteration_times", "pddp_final_max_defect", "pddp_set_graphs", "pddp_last_graph_launch_count", "pddp_alpha_shard_unique_id", "pddp_alpha_shard_init", "pddp_alpha_shard_stats", "pddp_set_bp_shape",
           "pddp_hardware_controls", "pddp_traj_f_encoded_size", "pddp_traj_f_encode", "pddp_traj_f_decode", "pddp_traj_f_pack_reference"]

_lib = None
FP = C.POINTER(C.c_float)
IP = C.POINTER(C.c_int)
DP = C.POINTER(C.c_double)


def load_library():
    """dlopen libpddp.so (built in-tree by parallel-ddp_b200/build.sh).  Raises PddpError if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PddpError(f"{LIB_PATH} is missing: run parallel-ddp_b200/build.sh (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    L.pddp_default_config_kuka.argtypes = [C.POINTER(Config), C.c_int, C.c_int]; L.pddp_default_config_kuka.restype = None
    L.pddp_create.argtypes = [C.POINTER(Config), C.POINTER(H)]
    L.pddp_destroy.argtypes = [H]; L.pddp_destroy.restype = None
    L.pddp_last_error.argtypes = [H]; L.pddp_last_error.restype = C.c_char_p
    L.pddp_solve.argtypes = [H, FP, FP, FP, C.c_int, C.c_int, C.c_int, FP, FP, FP, IP, IP, DP]
    L.pddp_solve_device.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, DP]
    L.pddp_make_inputs_kuka.argtypes = [C.c_int, C.c_int, C.c_uint, FP, FP, FP]
    L.pddp_unit_dynamics.argtypes = [H, FP, FP, C.c_int, FP]
    L.pddp_unit_integrator_gradient.argtypes = [H, FP, FP, C.c_int, FP, FP]
    L.pddp_set_array.argtypes = [H, C.c_char_p, C.c_void_p, C.c_long]
    L.pddp_get_array.argtypes = [H, C.c_char_p, C.c_void_p, C.c_long]
    L.pddp_phase_load_init.argtypes = [H, FP, FP, FP, C.c_int]
    for f in ("pddp_phase_backward_pass", "pddp_phase_forward_sweep", "pddp_phase_forward_sim", "pddp_phase_line_search", "pddp_phase_next_iteration"):
        getattr(L, f).argtypes = [H]
    L.pddp_last_phase_stats.argtypes = [H, DP, IP]
    L.pddp_last_launch_count.argtypes = [H]; L.pddp_last_launch_count.restype = C.c_long
    L.pddp_set_groups.argtypes = [H, C.c_int]
    L.pddp_selftest_rcp.argtypes = [C.POINTER(C.c_ulonglong)]
    L.pddp_selftest_sincos.argtypes = [C.POINTER(C.c_ulonglong)]
    L.pddp_set_warm_start.argtypes = [H, FP, FP, FP, FP]
    L.pddp_set_start_mode.argtypes = [H, C.c_int, C.c_int]
    L.pddp_mpc_init.argtypes = [H, FP, FP]
    L.pddp_set_skip_unchanged.argtypes = [H, C.c_int]
    L.pddp_set_x_target.argtypes = [H, FP]
    L.pddp_mpc_set_cost_shift.argtypes = [H, C.c_int]
    L.pddp_hardware_controls.argtypes = [C.c_int, C.c_double, FP, FP, FP, C.c_double, DP, DP, C.c_double, C.c_int, C.c_int, DP, C.c_double, DP, DP]
    L.pddp_traj_f_encoded_size.argtypes = [C.c_int, C.c_int, C.c_int]; L.pddp_traj_f_encoded_size.restype = C.c_long
    L.pddp_traj_f_encode.argtypes = [C.c_longlong, FP, C.c_int, FP, C.c_int, FP, C.c_int, C.c_void_p, C.c_long]; L.pddp_traj_f_encode.restype = C.c_long
    L.pddp_traj_f_decode.argtypes = [C.c_void_p, C.c_long, C.POINTER(C.c_longlong), IP, IP, IP, FP, FP, FP, C.c_long, C.c_long, C.c_long]; L.pddp_traj_f_decode.restype = C.c_long
    L.pddp_traj_f_pack_reference.argtypes = [C.c_longlong, FP, FP, FP, C.c_int, C.c_int, C.c_void_p, C.c_long]; L.pddp_traj_f_pack_reference.restype = C.c_long
    L.pddp_mpc_step.argtypes = [H, FP, FP, IP, C.c_int, C.c_int, C.c_int, FP, FP, FP, FP, IP, IP, IP]
    L.pddp_default_config.argtypes = [C.POINTER(Config), C.c_int, C.c_int, C.c_int]
    L.pddp_plant_dims.argtypes = [C.c_int, IP, IP, IP]
    L.pddp_register_plant.argtypes = [C.c_void_p]
    L.pddp_load_plant_library.argtypes = [C.c_char_p]
    L.pddp_plant_error.argtypes = []; L.pddp_plant_error.restype = C.c_char_p
    L.pddp_make_inputs.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint, FP, FP, FP]
    L.pddp_last_iteration_times.argtypes = [H, DP, DP, DP, DP, C.c_int]
    L.pddp_final_max_defect.argtypes = [H, FP]
    L.pddp_set_graphs.argtypes = [H, C.c_int, C.c_int]
    L.pddp_alpha_shard_unique_id.argtypes = [C.c_void_p]
    L.pddp_alpha_shard_init.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    L.pddp_alpha_shard_stats.argtypes = [H, DP, IP, IP]
    L.pddp_set_bp_shape.argtypes = [H, C.c_int]
    L.pddp_last_graph_launch_count.argtypes = [H]; L.pddp_last_graph_launch_count.restype = C.c_long
    L.pddp_unit_integrator.argtypes = [H, FP, FP, C.c_int, FP]
    L.pddp_unit_cost.argtypes = [H, FP, FP, FP, IP, C.c_int, FP, FP, FP]
    _lib = L
    return L


def default_config_kuka(N=128, batch=1, **over):
    L = load_library()
    c = Config()
    L.pddp_default_config_kuka(C.byref(c), N, batch)
    for k, v in over.items():
        setattr(c, k, v)
    return c


def default_config(plant, N, batch=1, **over):
    """pddp_default_config: the reference's compile-time defaults of a plant (config.cuh:21-136) as a run-time Config."""
    L = load_library()
    c = Config()
    if L.pddp_default_config(C.byref(c), plant, N, batch) != 0:
        raise PddpError(f"no defaults for PLANT {plant}")
    for k, v in over.items():
        setattr(c, k, v)
    return c


def plant_dims(plant):
    """(NUM_POS, STATE_SIZE, CONTROL_SIZE) of a built-in or registered plant."""
    L = load_library(); a, b, c = C.c_int(), C.c_int(), C.c_int()
    if L.pddp_plant_dims(plant, C.byref(a), C.byref(b), C.byref(c)) != 0:
        raise PddpError(f"unknown PLANT {plant}")
    return a.value, b.value, c.value


def load_plant_library(path):
    """Register a plant plug-in library (include/pddp_plant.h); returns its PLANT number."""
    L = load_library()
    rc = L.pddp_load_plant_library(os.fsencode(path))
    if rc < 1:
        raise PddpError(f"pddp_load_plant_library({path}): {L.pddp_plant_error().decode()}")
    return rc


def make_inputs(plant, N, batch, seed0=0):
    """Initial guess and goal of the reference's example for any built-in plant (WAFR_iLQR_examples.cu:19-33,67-121)."""
    L = load_library(); _, n, m = plant_dims(plant)
    x0 = np.zeros((batch, N, n), np.float32); u0 = np.zeros((batch, N, m), np.float32); xg = np.zeros((batch, n), np.float32)
    if L.pddp_make_inputs(plant, N, batch, seed0, x0.ctypes.data_as(FP), u0.ctypes.data_as(FP), xg.ctypes.data_as(FP)) != 0:
        raise PddpError(f"no example inputs for PLANT {plant}")
    return x0, u0, xg


def alpha_shard_unique_id():
    """128-byte NCCL unique id for Solver.alpha_shard_init (call on rank 0, hand to all ranks)"""
    L = load_library(); buf = (C.c_ubyte * 128)()
    if L.pddp_alpha_shard_unique_id(buf) != 0:
        raise PddpError("pddp_alpha_shard_unique_id: libnccl.so.2 could not be loaded")
    return bytes(buf)


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(FP)


def make_inputs_kuka(N, batch, seed0=0):
    """Synthetic benchmark inputs (WAFR_iLQR_examples.cu:67-121), one std::default_random_engine(seed) per problem."""
    L = load_library()
    x0 = np.zeros((batch, N, 14), np.float32); u0 = np.zeros((batch, N, 7), np.float32); xg = np.zeros((batch, 14), np.float32)
    L.pddp_make_inputs_kuka(N, batch, seed0, x0.ctypes.data_as(FP), u0.ctypes.data_as(FP), xg.ctypes.data_as(FP))
    return x0, u0, xg


def traj_f_pack_reference(utime, x, u, KT, steps, with_feedback=True):
    """Bytes of the lcmt_trajectory_f message the reference's MPC loop publishes for one arm (LCMHelpers.cuh:245-252)."""
    L = load_library()
    x, px = _f(x); u, pu = _f(u); KT, pk = _f(KT)
    need = L.pddp_traj_f_pack_reference(utime, px, pu, pk, steps, int(with_feedback), None, 0)
    if need < 0:
        raise PddpError("pddp_traj_f_pack_reference: bad arguments")
    buf = (C.c_ubyte * need)()
    L.pddp_traj_f_pack_reference(utime, px, pu, pk, steps, int(with_feedback), buf, need)
    return bytes(buf)


def traj_f_decode(data):
    """(utime, x, u, KT) of an encoded lcmt_trajectory_f message."""
    L = load_library(); n = len(data); raw = (C.c_ubyte * n).from_buffer_copy(data)
    t = C.c_longlong(0); xs = C.c_int(0); us = C.c_int(0); ks = C.c_int(0)
    if L.pddp_traj_f_decode(raw, n, C.byref(t), C.byref(xs), C.byref(us), C.byref(ks), None, None, None, 0, 0, 0) < 0:
        raise PddpError("not an lcmt_trajectory_f message")
    x = np.zeros(xs.value, np.float32); u = np.zeros(us.value, np.float32); KT = np.zeros(ks.value, np.float32)
    L.pddp_traj_f_decode(raw, n, None, None, None, None, x.ctypes.data_as(FP), u.ctypes.data_as(FP), KT.ctypes.data_as(FP), x.size, u.size, KT.size)
    return int(t.value), x, u, KT


def selftest_rcp():
    """Number of float bit patterns on which the library's reciprocal differs from IEEE 1.0f/x (must be 0)."""
    L = load_library(); bad = C.c_ulonglong(0)
    rc = L.pddp_selftest_rcp(C.byref(bad))
    if rc != 0:
        raise PddpError(f"pddp_selftest_rcp failed ({rc})")
    return int(bad.value)


def selftest_sincos():
    """Number of float bit patterns on which the dynamics' sin / cos differ from the CUDA library's sinf / cosf (must be 0)."""
    L = load_library(); bad = C.c_ulonglong(0)
    rc = L.pddp_selftest_sincos(C.byref(bad))
    if rc != 0:
        raise PddpError(f"pddp_selftest_sincos failed ({rc})")
    return int(bad.value)


class Solver:
    """Owns every device array of `batch` problems (what allocateMemory_GPU hands back as ~50 raw pointers)."""
    _SHAPES = None

    def __init__(self, cfg):
        self.L = load_library()
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.L.pddp_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise PddpError(f"pddp_create failed ({rc}): {self.L.pddp_last_error(None).decode()}")
        B, N, A, M = cfg.batch, cfg.N, cfg.n_alpha, cfg.M
        npos, n, m = plant_dims(cfg.plant)
        nm = n + m
        self.n, self.m, self.npos = n, m, npos
        f, i = np.float32, np.int32
        self.shapes = dict(x=((B, A, N, n), f), u=((B, A, N, m), f), d=((B, A, N, n), f), xp=((B, N, n), f), xp2=((B, N, n), f),
                           up=((B, N, m), f), dp=((B, N, n), f), AB=((B, N, nm, n), f), H=((B, N, nm, nm), f), g=((B, N, nm), f),
                           P=((B, N, n, n), f), p=((B, N, n), f), Pp=((B, N, n, n), f), pp=((B, N, n), f), KT=((B, N, m, n), f),
                           du=((B, N, m), f), ApBK=((B, N, n, n), f), Bdu=((B, N, n), f), xGoal=((B, n), f), costk=((B, A, N), f),
                           J=((B, A), f), dT=((B, A), f), dJexp=((B, 2 * M), f), rho=((B,), f), drho=((B,), f), prevJ=((B,), f),
                           dJ=((B,), f), z=((B,), f), iter=((B,), i), alphaIndex=((B,), i), ignore_defect=((B,), i), done=((B,), i),
                           accepted=((B,), i), final_src=((B,), i), dbg=((4096,), np.int64), Jout=((B, cfg.max_iter + 1), f), alphaOut=((B, cfg.max_iter + 1), i))

    def _ck(self, rc, what):
        if rc != 0:
            raise PddpError(f"{what} failed ({rc}): {self.L.pddp_last_error(self.h).decode()}")

    # ---- solver entry point -------------------------------------------------------------------------------------
    # ---- receding horizon (runiLQR_MPC_GPU) -------------------------------------------------------------------------
    def mpc_init(self, x_init, u_init):
        """Start plan of the receding-horizon loop: x_init [B,N,14], u_init [B,N,7]; they also become the published plan."""
        B, N = self.cfg.batch, self.cfg.N
        self.mpc_x, px = _f(np.broadcast_to(x_init, (B, N, self.n)).copy()); self.mpc_u, pu = _f(np.broadcast_to(u_init, (B, N, self.m)).copy())
        self.mpc_KT = np.zeros((B, N, self.n*self.m), np.float32)
        self._ck(self.L.pddp_mpc_init(self.h, px, pu), "pddp_mpc_init")

    def mpc_step(self, xActual, xGoal, shiftAmount, max_iter, clear_vars=0, ignoreFirstDefectFlag=0):
        """One runiLQR_MPC_GPU call per problem.  Returns dict(x, u, KT (the published plan, updated in place), Jout, alphaOut, iters,
        last_successful_solve)."""
        B = self.cfg.batch; L1 = self.cfg.max_iter + 1
        xa, pxa = _f(np.broadcast_to(xActual, (B, self.n))); xg, pg = _f(np.broadcast_to(xGoal, (B, self.n)))
        sh = np.ascontiguousarray(np.broadcast_to(shiftAmount, (B,)), dtype=np.int32)
        Jout = np.empty((B, L1), np.float32); aOut = np.empty((B, L1), np.int32); iters = np.empty(B, np.int32); lss = np.empty(B, np.int32)
        rc = self.L.pddp_mpc_step(self.h, pxa, pg, sh.ctypes.data_as(IP), max_iter, clear_vars, ignoreFirstDefectFlag,
                                  self.mpc_x.ctypes.data_as(FP), self.mpc_u.ctypes.data_as(FP), self.mpc_KT.ctypes.data_as(FP),
                                  Jout.ctypes.data_as(FP), aOut.ctypes.data_as(IP), iters.ctypes.data_as(IP), lss.ctypes.data_as(IP))
        self._ck(rc, "pddp_mpc_step")
        return dict(x=self.mpc_x, u=self.mpc_u, KT=self.mpc_KT, Jout=Jout, alphaOut=aOut, iters=iters, last_successful_solve=lss)

    def set_warm_start(self, KT0, P0, p0, d0):
        """runiLQR_GPU's KT0, P0, p0, d0: [B,N,98], [B,N,196], [B,N,14], [B,N,14] in the reference layouts."""
        B, N = self.cfg.batch, self.cfg.N
        a, pa = _f(np.broadcast_to(KT0, (B, N, self.n*self.m))); b, pb = _f(np.broadcast_to(P0, (B, N, self.n*self.n)))
        c, pc = _f(np.broadcast_to(p0, (B, N, self.n))); d, pd = _f(np.broadcast_to(d0, (B, N, self.n)))
        self._ck(self.L.pddp_set_warm_start(self.h, pa, pb, pc, pd), "pddp_set_warm_start")

    def runiLQR_GPU(self, x0, u0, xGoal, forwardRolloutFlag=0, clearVarsFlag=1, ignoreFirstDefectFlag=1, want_times=False,
                    KT0=None, P0=None, p0=None, d0=None):
        """Batched runiLQR_GPU.  x0 [B,N,14], u0 [B,N,7], xGoal [B,14] (or [14]).  Returns dict(x, u, Jout, alphaOut, iters[, times_ms]).
        KT0, P0, p0, d0 (reference argument order: x0, u0, KT0, P0, p0, d0) are used when clearVarsFlag = 0."""
        B, N = self.cfg.batch, self.cfg.N
        if KT0 is not None:
            self.set_warm_start(KT0, P0, p0, d0)
        x0, px = _f(np.broadcast_to(x0, (B, N, self.n))); u0, pu = _f(np.broadcast_to(u0, (B, N, self.m))); xg, pg = _f(np.broadcast_to(xGoal, (B, self.n)))
        L1 = self.cfg.max_iter + 1
        x = np.empty((B, N, self.n), np.float32); u = np.empty((B, N, self.m), np.float32)
        Jout = np.empty((B, L1), np.float32); aOut = np.empty((B, L1), np.int32); iters = np.empty(B, np.int32)
        times = np.zeros(6, np.float64)
        rc = self.L.pddp_solve(self.h, px, pu, pg, forwardRolloutFlag, clearVarsFlag, ignoreFirstDefectFlag,
                               x.ctypes.data_as(FP), u.ctypes.data_as(FP), Jout.ctypes.data_as(FP), aOut.ctypes.data_as(IP),
                               iters.ctypes.data_as(IP), times.ctypes.data_as(DP) if want_times else None)
        self._ck(rc, "pddp_solve")
        out = dict(x=x, u=u, Jout=Jout, alphaOut=aOut, iters=iters)
        if want_times:
            out["times_ms"] = dict(zip(("total", "sim", "sweep", "bp", "nis", "init"), times))
        return out

    def solve_device(self, d_x0, d_u0, d_xg, d_x, d_u, d_J, d_a, d_it, ignoreFirstDefectFlag=1, times=None):
        """Device-resident variant: arguments are raw device pointers (ints)."""
        rc = self.L.pddp_solve_device(self.h, d_x0, d_u0, d_xg, ignoreFirstDefectFlag, d_x, d_u, d_J, d_a, d_it,
                                      times.ctypes.data_as(DP) if times is not None else None)
        self._ck(rc, "pddp_solve_device")

    def set_x_target(self, xTarget):
        """EE_COST: xTarget [B,14] of the nominal-state cost terms (runiLQR_MPC_GPU's gv->xTarget), or None."""
        if xTarget is None:
            self._ck(self.L.pddp_set_x_target(self.h, None), "pddp_set_x_target"); return
        a, pa = _f(np.broadcast_to(np.asarray(xTarget, np.float32).reshape(-1, self.n), (self.cfg.batch, self.n)))
        self._ck(self.L.pddp_set_x_target(self.h, pa), "pddp_set_x_target")

    def mpc_set_cost_shift(self, on):
        """runiLQR_MPC_GPU's use_cost_shift: final pose weights on the last shiftAmount+1 knots of every receding-horizon step (EE_COST)."""
        self._ck(self.L.pddp_mpc_set_cost_shift(self.h, int(on)), "pddp_mpc_set_cost_shift")

    def set_skip_unchanged(self, on):
        """Opt-in: no gradient refresh for problems whose line search was rejected (results unchanged)."""
        self._ck(self.L.pddp_set_skip_unchanged(self.h, int(on)), "pddp_set_skip_unchanged")

    def set_groups(self, groups):
        """Problem groups iterated on separate streams (overlap of latency- and throughput-bound kernels); returns the value in effect."""
        return int(self.L.pddp_set_groups(self.h, groups))

    def iteration_times(self):
        """per-iteration device times (ms) of the last timed solve run as one problem group: dict(sim, sweep, bp, nis) of arrays"""
        cap = self.cfg.max_iter; a = [np.zeros(cap, np.float64) for _ in range(4)]
        cnt = self.L.pddp_last_iteration_times(self.h, *[v.ctypes.data_as(DP) for v in a], cap)
        return dict(zip(("sim", "sweep", "bp", "nis"), [v[:cnt] for v in a]))

    def final_max_defect(self):
        d = np.zeros(self.cfg.batch, np.float32)
        self._ck(self.L.pddp_final_max_defect(self.h, d.ctypes.data_as(FP)), "pddp_final_max_defect")
        return d

    def set_bp_shape(self, shape):
        """backward-pass shape: 0 by launch size, 1 warp chains, 2 block-cooperative (identical results)"""
        self._ck(self.L.pddp_set_bp_shape(self.h, int(shape)), "pddp_set_bp_shape")

    def alpha_shard_init(self, rank, nranks, unique_id):
        """shard the line search's step sizes over `nranks` GPUs (one process each); unique_id: the 128 bytes of alpha_shard_unique_id()
        made on rank 0 and handed to every rank"""
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.L.pddp_alpha_shard_init(self.h, int(rank), int(nranks), buf), "pddp_alpha_shard_init")

    def alpha_shard_stats(self):
        us = C.c_double(); a0 = C.c_int(); cnt = C.c_int()
        self.L.pddp_alpha_shard_stats(self.h, C.byref(us), C.byref(a0), C.byref(cnt))
        return dict(exchange_us_per_iteration=us.value, a_first=a0.value, a_cnt=cnt.value)

    def set_graphs(self, on, iterations_per_graph=0):
        """CUDA-graph replay of the iteration loop (default on, 10 iterations per graph)"""
        self._ck(self.L.pddp_set_graphs(self.h, int(on), int(iterations_per_graph)), "pddp_set_graphs")

    def graph_launch_count(self):
        return int(self.L.pddp_last_graph_launch_count(self.h))

    def launch_count(self):
        return int(self.L.pddp_last_launch_count(self.h))

    # ---- plant plug-ins -----------------------------------------------------------------------------------------
    def dynamics(self, x, u):
        x, px = _f(x); u, pu = _f(u); n = x.shape[0]
        qdd = np.empty((n, self.npos), np.float32)
        self._ck(self.L.pddp_unit_dynamics(self.h, px, pu, n, qdd.ctypes.data_as(FP)), "pddp_unit_dynamics")
        return qdd

    def integratorGradient(self, x, u):
        x, px = _f(x); u, pu = _f(u); n = x.shape[0]
        AB = np.empty((n, self.n + self.m, self.n), np.float32); qdd = np.empty((n, self.npos), np.float32)
        self._ck(self.L.pddp_unit_integrator_gradient(self.h, px, pu, n, AB.ctypes.data_as(FP), qdd.ctypes.data_as(FP)), "pddp_unit_integrator_gradient")
        return AB, qdd

    def integrator(self, x, u):
        """x_{k+1} = _integrator(x_k, u_k) of the configured INTEGRATOR (plug-in plants)"""
        x, px = _f(x); u, pu = _f(u); n = x.shape[0]
        xn = np.empty((n, self.n), np.float32)
        self._ck(self.L.pddp_unit_integrator(self.h, px, pu, n, xn.ctypes.data_as(FP)), "pddp_unit_integrator")
        return xn

    def cost(self, x, u, xGoal, knot):
        """costFunc and costGrad of a plug-in plant on n samples: returns (J [n], H [n, n+m, n+m], g [n, n+m])"""
        x, px = _f(x); u, pu = _f(u); xg, pg = _f(xGoal); n = x.shape[0]; nm = self.n + self.m
        kn = np.ascontiguousarray(np.broadcast_to(knot, (n,)), dtype=np.int32)
        J = np.empty(n, np.float32); H = np.empty((n, nm, nm), np.float32); g = np.empty((n, nm), np.float32)
        self._ck(self.L.pddp_unit_cost(self.h, px, pu, pg, kn.ctypes.data_as(IP), n, J.ctypes.data_as(FP), H.ctypes.data_as(FP), g.ctypes.data_as(FP)), "pddp_unit_cost")
        return J, H, g

    # ---- phase-level access -------------------------------------------------------------------------------------
    def get(self, name):
        shp, dt = self.shapes[name]
        a = np.empty(shp, dt)
        self._ck(self.L.pddp_get_array(self.h, name.encode(), a.ctypes.data_as(C.c_void_p), a.nbytes), f"get {name}")
        return a

    def set(self, name, value):
        shp, dt = self.shapes[name]
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=dt), shp))
        self._ck(self.L.pddp_set_array(self.h, name.encode(), a.ctypes.data_as(C.c_void_p), a.nbytes), f"set {name}")

    def load_init(self, x0, u0, xGoal, ignoreFirstDefectFlag=1):
        B, N = self.cfg.batch, self.cfg.N
        x0, px = _f(np.broadcast_to(x0, (B, N, self.n))); u0, pu = _f(np.broadcast_to(u0, (B, N, self.m))); xg, pg = _f(np.broadcast_to(xGoal, (B, self.n)))
        self._ck(self.L.pddp_phase_load_init(self.h, px, pu, pg, ignoreFirstDefectFlag), "load_init")

    def backwardPassGPU(self):
        self._ck(self.L.pddp_phase_backward_pass(self.h), "backward_pass")

    def forwardSweep(self):
        self._ck(self.L.pddp_phase_forward_sweep(self.h), "forward_sweep")

    def forwardSimGPU(self):
        self._ck(self.L.pddp_phase_forward_sim(self.h), "forward_sim")
        self._ck(self.L.pddp_phase_line_search(self.h), "line_search")

    def forwardSimOnly(self):
        self._ck(self.L.pddp_phase_forward_sim(self.h), "forward_sim")

    def lineSearchAcceptReject(self):
        self._ck(self.L.pddp_phase_line_search(self.h), "line_search")

    def nextIterationSetupGPU(self):
        self._ck(self.L.pddp_phase_next_iteration(self.h), "next_iteration")

    def last_phase_ms(self):
        ms = C.c_double(); n = C.c_int()
        self.L.pddp_last_phase_stats(self.h, C.byref(ms), C.byref(n))
        return ms.value, n.value

    def freeMemory_GPU(self):
        if self.h:
            self.L.pddp_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.freeMemory_GPU()
        except Exception:
            pass


def hardware_controls(x, u, KT, t0, qActual, qdActual, tActual, time_step, use_feedback=True, pd_gains_on_state=False, u_prev=None, alpha=0.0):
    """getHardwareControls (MPCHelpers.cuh:817-858): joint command from the published plan x [N,14], u [N,7], KT [N,98] and the
    measured state at time tActual (microseconds, like t0).  Returns (err, q_out[7], u_out[7]); err = 1 when tActual is beyond
    the plan.  u_prev (float64[7], updated in place) with alpha > 0 turns on the reference's exponential smoothing."""
    L = load_library()
    x, px = _f(x); u, pu = _f(u); KT, pk = _f(KT)
    qa = np.ascontiguousarray(qActual, np.float64); qda = np.ascontiguousarray(qdActual, np.float64)
    qo = np.zeros(7, np.float64); uo = np.zeros(7, np.float64)
    pp = u_prev.ctypes.data_as(DP) if u_prev is not None else None
    err = L.pddp_hardware_controls(x.size // 14, float(time_step), px, pu, pk, float(t0), qa.ctypes.data_as(DP), qda.ctypes.data_as(DP), float(tActual),
                                   int(use_feedback), int(pd_gains_on_state), pp, float(alpha), qo.ctypes.data_as(DP), uo.ctypes.data_as(DP))
    return err, qo, uo
