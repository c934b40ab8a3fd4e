"""End-effector cost: the CUDA path against the reference's GPU dump (gpurun_out/golden_raw/ee_solve_*.bin, written by
tests/golden/make_goldens.py gpu ee_) and against the oracle's GPU-arithmetic build.  Development aid; the test proper is
tests/test_gpu_parity.py::test_ee_*."""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refdump, oracle_lib as ol
pddp = importlib.import_module("parallel-ddp_b200")

raw = os.path.join(ROOT, "gpurun_out", "golden_raw")
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(raw, "ee_solve_G_N32_s0-3_tol0.bin")
d = refdump.load(src) if src.endswith(".bin") else dict(np.load(src))
N, A, M, ns = [int(v) for v in d["meta"]]
W = [float(v) for v in d["weights"]]
x0 = d["x_in"].reshape(ns, N, 14); u0 = d["u_in"].reshape(ns, N, 7)
xg = np.zeros((ns, 14), np.float32); xg[:, :6] = d["xGoal"]
L1 = 101
Jr = d["Jout"].reshape(ns, L1); ar = d["alphaOut"].reshape(ns, L1)
cfg = pddp.default_config_kuka(N, ns, ee_cost=1, **dict(zip(pddp.EE_WEIGHT_NAMES, W)))
s = pddp.Solver(cfg)
o = s.runiLQR_GPU(x0, u0, xg)
print("reference iters", d["iters"], "cuda iters", o["iters"])
for b in range(ns):
    same_a = np.array_equal(ar[b], o["alphaOut"][b]); same_J = Jr[b].tobytes() == o["Jout"][b].tobytes()
    print(f"problem {b}: alpha trace equal {same_a}, cost trace equal {same_J}, x equal {np.array_equal(d['x_out'].reshape(ns,N,14)[b], o['x'][b])}, u equal {np.array_equal(d['u_out'].reshape(ns,N,7)[b], o['u'][b])}")
    if not (same_a and same_J):
        k = int(np.argmax((ar[b] != o["alphaOut"][b]) | (Jr[b].view(np.int32) != o["Jout"][b].view(np.int32))))
        print("   first difference at iteration", k, "ref", ar[b][max(0,k-1):k+3], Jr[b][max(0,k-1):k+3], "cuda", o["alphaOut"][b][max(0,k-1):k+3], o["Jout"][b][max(0,k-1):k+3])
# oracle (GPU arithmetic)
Lf = ol.lib(True); c = ol.kuka_cfg(N, True, ee_weights=W)
for b in range(min(ns, 2)):
    xo = np.zeros((N, 14), np.float32); uo = np.zeros((N, 7), np.float32); Jo = np.full(L1, np.nan, np.float32); ao = np.full(L1, -99, np.int32)
    Lf.orc_solve(C.byref(c), ol.fptr(np.ascontiguousarray(x0[b])), ol.fptr(np.ascontiguousarray(u0[b])), ol.fptr(np.ascontiguousarray(xg[b])), ol.fptr(xo), ol.fptr(uo), ol.fptr(Jo), ol.iptr(ao))
    print(f"oracle problem {b}: alpha equal ref {np.array_equal(ao, ar[b])}, cost equal ref {Jo.tobytes() == Jr[b].tobytes()}; equal cuda {np.array_equal(ao, o['alphaOut'][b])} {Jo.tobytes() == o['Jout'][b].tobytes()}")
    if not np.array_equal(ao, ar[b]) or Jo.tobytes() != Jr[b].tobytes():
        k = int(np.argmax((ao != ar[b]) | (Jo.view(np.int32) != Jr[b].view(np.int32))))
        print("   first difference at iteration", k, "ref", ar[b][max(0,k-1):k+3], Jr[b][max(0,k-1):k+3], "oracle", ao[max(0,k-1):k+3], Jo[max(0,k-1):k+3])
