"""Tiny end-to-end run of every kernel (used under compute-sanitizer)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pddp = importlib.import_module("parallel-ddp_b200")
N, B = 32, 2
x0, u0, xg = pddp.make_inputs_kuka(N, B, 0)
s = pddp.Solver(pddp.default_config_kuka(N, B, max_iter=2))
o = s.runiLQR_GPU(x0, u0, xg)
print("iters", o["iters"], "alpha", o["alphaOut"][:, :3], "J", o["Jout"][:, :3])
q = s.dynamics(x0[0, :4], u0[0, :4]); AB, _ = s.integratorGradient(x0[0, :2], u0[0, :2])
print("qdd", q[0], "AB00", AB[0, 0, :3])
