"""Receding-horizon step latency: python tools/mpc_time.py [batch] [max_iter] [shift]  (wall time of pddp_mpc_step, host buffers)"""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pddp = importlib.import_module("parallel-ddp_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 5
shift = int(sys.argv[3]) if len(sys.argv) > 3 else 2
N = 128
x0, u0, xg = pddp.make_inputs_kuka(N, B, 0)
u0 = np.zeros_like(u0) + 0.01
s = pddp.Solver(pddp.default_config_kuka(N, B, tol_cost=1e-4, gravity=0.0, max_iter=max(max_iter, 8)))
s.mpc_init(x0, u0)
rng = np.random.default_rng(0); ts = []
for st in range(25):
    sh = 0 if st == 0 else shift
    xa = np.ascontiguousarray(s.mpc_x[:, sh] + 0.002 * rng.standard_normal((B, 14)), np.float32)
    t0 = time.perf_counter(); o = s.mpc_step(xa, xg, sh, max_iter, clear_vars=1 if st == 0 else 0); ts.append(time.perf_counter() - t0)
ts = np.array(ts[5:]) * 1e3
out = dict(batch=B, N=N, max_iter=max_iter, shift=shift, step_ms_median=float(np.median(ts)), step_ms_p90=float(np.percentile(ts, 90)),
           iters_per_step=float(np.mean(o["iters"])), arms_steps_per_s=float(B / (np.median(ts) * 1e-3)), launches=s.launch_count())
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True); json.dump(out, open(f"gpurun_out/mpc_time_b{B}.json", "w"))
