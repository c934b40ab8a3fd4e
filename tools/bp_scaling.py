"""Backward-pass kernel alone at growing batch: algorithmic GB/s against the measured HBM peak (python tools/bp_scaling.py [B...])"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pddp = importlib.import_module("parallel-ddp_b200")
if os.environ.get("PDDP_LIB"):
    pddp.LIB_PATH = os.path.join(ROOT, os.environ["PDDP_LIB"])
BYTES = 656452
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
N = 128
rows = []
for B in [int(a) for a in sys.argv[1:]] or [64, 256, 512, 1024, 2048, 4096]:
    x0, u0, xg = pddp.make_inputs_kuka(N, min(B, 64), 0)
    reps = (B + 63) // 64
    x0 = np.tile(x0, (reps, 1, 1))[:B]; u0 = np.tile(u0, (reps, 1, 1))[:B]; xg = np.tile(xg, (reps, 1))[:B]
    s = pddp.Solver(pddp.default_config_kuka(N, B, max_iter=3))
    s.load_init(x0, u0, xg)
    ts = []
    for it in range(6):
        s.backwardPassGPU(); ms, _ = s.last_phase_ms(); ts.append(ms)
        if it < 2:   # two real iterations so that P,p are populated, then repeat the pass on the same state
            s.forwardSweep(); s.forwardSimGPU(); s.nextIterationSetupGPU()
    t = float(np.median(ts[2:]))
    gbs = B * BYTES / (t * 1e-3) / 1e9
    rows.append(dict(batch=B, bp_ms=t, algorithmic_GBps=gbs, frac_of_measured_hbm_peak=gbs / peak))
    print(rows[-1] if not os.environ.get('PDDP_LIB') else (B, round(t*1e3, 1)), flush=True)
    s.freeMemory_GPU()
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "bp_scaling.json"), "w"), indent=1)
