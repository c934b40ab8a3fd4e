"""Short run of the headline configuration for profilers: python tools/prof_run.py [max_iter] [batch]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pddp = importlib.import_module("parallel-ddp_b200")
if os.environ.get("PDDP_LIB"):
    pddp.LIB_PATH = os.path.abspath(os.environ["PDDP_LIB"])
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N = 128
x0, u0, xg = pddp.make_inputs_kuka(N, B, 0)
kw = {}
if os.environ.get("PDDP_EE"):      # end-effector cost with the reference's default weights and the example's goal pose
    kw["ee_cost"] = 1; xg[:] = 0; xg[:, :6] = (0.3638, 0.0, 1.0628, 0.5*3.14159, 0.0, 0.5*3.14159)
s = pddp.Solver(pddp.default_config_kuka(N, B, max_iter=iters, **kw))
o = s.runiLQR_GPU(x0, u0, xg, want_times=True)
print({k: round(v, 3) for k, v in o["times_ms"].items()}, "launches", s.launch_count())
