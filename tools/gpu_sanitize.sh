#!/bin/bash
# compute-sanitizer over one small solve (N=32, 2 problems, 3 iterations), a warm start and two receding-horizon steps, with the
# joint-space cost and with the end-effector cost (xTarget set)
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
sl = pddp.Solver(pddp.default_config_kuka(N, B, max_iter=3, use_limits=1)); ol_ = sl.runiLQR_GPU(x0, u0, xg)      # limit penalties (USE_LIMITS_FLAG 1)
print("done", o["iters"], o2["iters"], ol_["iters"])
# end-effector cost: cold solve, rollout start, receding horizon with xTarget
xe = np.zeros((B, 14), np.float32); xe[:, :6] = (0.3638, 0.0, 1.0628, 1.570795, 0.0, 1.570795)
w = dict(zip(pddp.EE_WEIGHT_NAMES, (0.1, 0.01, 1000.0, 10.0, 1e-4, 0.1, 1000.0, 1e-3, 1.0)))
e = pddp.Solver(pddp.default_config_kuka(N, B, max_iter=3, ee_cost=1, gravity=0.0, **w))
oe = e.runiLQR_GPU(x0, u0, xe)
oe2 = e.runiLQR_GPU(oe["x"], oe["u"], xe, forwardRolloutFlag=1, clearVarsFlag=0, KT0=z(B, N, 98), P0=z(B, N, 196), p0=z(B, N, 14), d0=z(B, N, 14))
e.set_x_target(xg)
e.mpc_init(x0, u0 * 0 + 0.01)
for st in range(3):
    e.mpc_step(e.mpc_x[:, 0].copy(), xe, 0 if st == 0 else 2, 2, clear_vars=1 if st == 0 else 0)
print("done ee", oe["iters"], oe2["iters"])
# round 2: both backward-pass shapes, graph replay off / on, plug-in plants with all three integrators incl. the rho retry and a
# receding-horizon step
for shape in (1, 2):
    s2 = pddp.Solver(pddp.default_config_kuka(N, B, max_iter=2)); s2.set_bp_shape(shape); s2.set_graphs(shape == 1, 2); s2.runiLQR_GPU(x0, u0, xg)
for plant, integ, A in ((1, 3, 1), (2, 2, 8), (3, 3, 16), (3, 1, 4)):
    c = pddp.default_config(plant, 32, B, n_alpha=A, integrator=integ, max_iter=3)
    if plant == 2: c.R = -0.5          # forces the rho retry of the 1-D inverse
    a0, b0, g0 = pddp.make_inputs(plant, 32, B, seed0=2)
    sp = pddp.Solver(c); op = sp.runiLQR_GPU(a0, b0, g0)
    sp.mpc_init(a0, b0); sp.mpc_step(a0[:, 1].copy(), g0, 1, 2)
    sp.dynamics(a0[0, :3], b0[0, :3]); sp.integratorGradient(a0[0, :3], b0[0, :3]); sp.integrator(a0[0, :3], b0[0, :3]); sp.cost(a0[0, :3], b0[0, :3], g0[0], [0, 5, 31])
    print("done plant", plant, integ, op["iters"])
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"; timeout 500 compute-sanitizer --tool $tool python /tmp/san.py > /tmp/san_$tool.txt 2>&1
  grep -E "^done|ERROR SUMMARY|RACECHECK SUMMARY" /tmp/san_$tool.txt; echo "hazard / error records: $(grep -cE "hazard detected|Invalid|Error:" /tmp/san_$tool.txt)"; grep -E "hazard detected|Invalid" /tmp/san_$tool.txt | sort | uniq -c | head -5
done
