import importlib, os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pddp = importlib.import_module("parallel-ddp_b200")
pddp.LIB_PATH = os.path.join(os.path.dirname(pddp.LIB_PATH), "libpddp_trace.so")
N, B = 128, int(sys.argv[1]) if len(sys.argv) > 1 else 64
x0, u0, xg = pddp.make_inputs_kuka(N, B, 0)
s = pddp.Solver(pddp.default_config_kuka(N, B, max_iter=3))
s.load_init(x0, u0, xg); s.backwardPassGPU(); s.forwardSweep(); s.forwardSimGPU(); s.nextIterationSetupGPU(); s.backwardPassGPU()
d = s.get("dbg")[:96].reshape(-1, 12)
for k in range(1, 7):
    r = d[k]; nxt = d[k+1][0]
    seg = [r[1]-r[0], r[2]-r[1], r[3]-r[2], r[4]-r[3], r[5]-r[4], r[6]-r[5], r[7]-r[6], r[8]-r[7], r[9]-r[8], nxt - r[9]]
    print("   E detail: compute", int(r[10]-r[7]), "stores", int(r[11]-r[10]), "rest", int(r[8]-r[11]))
    print("knot", k, "total", int(nxt - r[0]), dict(zip(["mbarwait", "A", "B1", "GJ", "B2bar", "C", "D", "E", "Ebar", "top"], map(int, seg))))
