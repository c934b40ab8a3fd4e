"""Headline leg of bench.py alone (device-timed, inputs resident in HBM) + un-overlapped phase times: python tools/quick_rate.py [steps] [batch] [groups]"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pddp = importlib.import_module("parallel-ddp_b200")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
G = int(sys.argv[3]) if len(sys.argv) > 3 else 4
N, L1 = 128, 101
solver = pddp.Solver(pddp.default_config_kuka(N, B, device=0, tol_cost=0.0))
x0, u0, xg = pddp.make_inputs_kuka(N, B, seed0=0)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(a).to(dev) for a in (x0, u0, xg)]
ox = torch.empty_like(d[0]); ou = torch.empty_like(d[1])
dJ = torch.empty((B, L1), dtype=torch.float32, device=dev); da = torch.empty((B, L1), dtype=torch.int32, device=dev); dit = torch.empty(B, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); times = np.zeros(6, np.float64)
def step():
    flush.fill_(1); torch.cuda.synchronize()
    solver.solve_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), ox.data_ptr(), ou.data_ptr(), dJ.data_ptr(), da.data_ptr(), dit.data_ptr(), 1, times)
    return times.copy()
solver.set_groups(G)
for _ in range(3): step()
ms = sum(step()[0] for _ in range(steps)); its = int(dit.sum().item())
print(f"groups {G}: {its*steps/(ms/1e3):.0f} iterations/s  ({ms/steps:.2f} ms per step, {1e3*ms/steps/100:.1f} us per iteration)")
solver.set_groups(1); step(); ph = sum(step() for _ in range(3)) / 3
print("groups 1, us per iteration: " + "  ".join(f"{k} {1e3*v/100:.1f}" for k, v in zip(("total", "sim+select", "sweep", "bp", "nis", "init+store"), ph)))
print("checksum", float(ox.double().sum().item()), float(dJ[:, -1].double().sum().item()))
