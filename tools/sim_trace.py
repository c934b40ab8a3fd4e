"""Phase latencies of ONE forward-dynamics evaluation on a lone warp (trace build, tools/build_trace.sh): python tools/sim_trace.py
The stamps come from the last evaluation block 0 ran: with one problem that is the simulation kernel of the last iteration, one warp per SM."""
import ctypes as C, importlib, os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pddp = importlib.import_module("parallel-ddp_b200")
pddp.LIB_PATH = os.path.join(os.path.dirname(pddp.LIB_PATH), "libpddp_trace.so")
N, B = 128, int(sys.argv[1]) if len(sys.argv) > 1 else 1
x0, u0, xg = pddp.make_inputs_kuka(N, B, 0)
s = pddp.Solver(pddp.default_config_kuka(N, B, max_iter=3))
s.runiLQR_GPU(x0, u0, xg)
lib = C.CDLL(pddp.LIB_PATH); out = (C.c_longlong * 16)()
assert lib.pddp_debug_simtrace(out) == 0
t = np.array(out[:10], np.int64); names = ["sin/cos + joint T", "T chain", "body block (TA, J, Iw)", "Icrbs suffix + twists", "crm J, JdotV prefix", "wrench + F", "M (28 dots)", "W suffix + tau", "Gauss-Jordan solve"]
for nme, d in zip(names, np.diff(t)): print(f"{nme:28s} {int(d):6d} cycles")
print(f"{'total':28s} {int(t[9]-t[0]):6d} cycles")
