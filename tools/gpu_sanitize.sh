#!/bin/bash
# compute-sanitizer over one small solve (N=32, 2 problems, 3 iterations), a warm start and two receding-horizon steps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import importlib, sys, os, numpy as np
sys.path.insert(0, os.getcwd())
pddp = importlib.import_module("parallel-ddp_b200")
N, B = 32, 2
x0, u0, xg = pddp.make_inputs_kuka(N, B, 1)
s = pddp.Solver(pddp.default_config_kuka(N, B, max_iter=3))
o = s.runiLQR_GPU(x0, u0, xg)
z = lambda *sh: np.zeros(sh, np.float32)
o2 = s.runiLQR_GPU(o["x"], o["u"], xg, forwardRolloutFlag=1, clearVarsFlag=0, KT0=z(B, N, 98), P0=z(B, N, 196), p0=z(B, N, 14), d0=z(B, N, 14))
s.mpc_init(x0, u0 * 0 + 0.01)
for st in range(2):
    s.mpc_step(s.mpc_x[:, 0].copy(), xg, 0 if st == 0 else 2, 2, clear_vars=1 if st == 0 else 0)
print("done", o["iters"], o2["iters"])
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"; timeout 500 compute-sanitizer --tool $tool python /tmp/san.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Error|hazard" | head -8
done
