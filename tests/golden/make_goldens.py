#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference (oracle/_ref/ref_driver_N*,
built by `make -C oracle ref` from /root/reference -- see oracle/Makefile).

  python tests/golden/make_goldens.py host      # here (no GPU): reference HOST instantiation  -> *_H_*.npz
  python tests/golden/make_goldens.py gpu [prefix]  # on the GPU box (under gpurun): raw dumps   -> gpurun_out/golden_raw/*.bin
  python tests/golden/make_goldens.py import    # here: gpurun_out/golden_raw/*.bin             -> *_G_*.npz

Fixtures are compressed .npz files holding the named arrays ref_driver wrote; redundant broadcast copies
(x1..x15 after `nis`/`init`) are dropped to keep them small."""
import os
import re
import subprocess
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refdump  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
RAW = os.path.join(ROOT, "gpurun_out", "golden_raw")
_DROP = re.compile(r"it\d+\.(nis|init)\.[xud]([1-9]|1\d)$")


def run(N, *args):
    exe = os.path.join(REF, f"ref_mpc_N{N}" if args and args[0] == "mpc" else f"ref_driver_N{N}")
    if isinstance(N, str):                              # pendulum / cart-pole / quadrotor builds: ref_p<plant>_i<integrator>_N<knots>_a<alphas>
        exe = os.path.join(REF, "ref_" + N)
    if args and args[0] == "mpc_ee":                    # receding horizon built with EE_COST 1 (ref_mpc.cu -DEE_COST=1)
        exe, args = os.path.join(REF, f"ref_mpc_ee_N{N}"), ("mpc",) + tuple(args[1:])
        if "CS" in args:                                # use_cost_shift = 1: the flag follows the output file on ref_mpc's command line
            args = tuple(a for a in args if a != "CS") + (1,)
    if args and str(args[0]).startswith("eelim_"):       # end-effector cost with the limit penalties (ref_ee.cu built with -DUSE_LIMITS_FLAG=1)
        exe, args = os.path.join(REF, f"ref_eelim_N{N}"), (args[0][6:],) + tuple(args[1:])
    if args and str(args[0]).startswith("eesa_"):        # end-effector cost with the smooth-abs pose term (ref_ee.cu built with -DUSE_SMOOTH_ABS=1)
        exe, args = os.path.join(REF, f"ref_eesa_N{N}"), (args[0][5:],) + tuple(args[1:])
    if args and str(args[0]).startswith("lim_"):         # joint-space cost with the limit penalties (ref_driver.cu built with -DUSE_LIMITS_FLAG=1)
        exe, args = os.path.join(REF, f"ref_lim_N{N}"), (args[0][4:],) + tuple(args[1:])
    if args and str(args[0]).startswith("ee_"):          # end-effector cost build (EE_COST 1): oracle/ref_harness/ref_ee.cu
        exe, args = os.path.join(REF, f"ref_ee_N{N}"), (args[0][3:],) + tuple(args[1:])
    print("+", exe, *args, flush=True)
    subprocess.run([exe, *[str(a) for a in args]], check=True, stdout=subprocess.DEVNULL)


def to_npz(binpath, npzpath):
    d = refdump.load(binpath)
    d = {k: v for k, v in d.items() if not _DROP.search(k)}
    np.savez_compressed(npzpath, **d)
    print("  ->", os.path.relpath(npzpath, ROOT), f"{os.path.getsize(npzpath)/1024:.0f} KiB")


def jobs(hw):
    """(N, args-without-outfile, name)"""
    t = "H" if hw == "H" else "G"
    out = [(32, ("unit", t, 64, 7), f"unit_{t}"),
           (32, ("trace", t, 0, 0.0, 2), f"trace_{t}_N32_s0_tol0"),
           (32, ("trace", t, 3, 0.0001, 1), f"trace_{t}_N32_s3_tol1e-4"),
           (128, ("trace", t, 0, 0.0, 1), f"trace_{t}_N128_s0_tol0"),
           # end-effector cost: cost / gradient / Hessian of costGradientHessianKern (G) / ...Threaded (H) on random states
           (32, ("ee_unit", t, 64, 7), f"ee_unit_{t}"),
           # USE_LIMITS_FLAG 1: cost / gradient on random states (host-evaluated in both dumps), every phase of the first iterations
           (32, ("eesa_unit", t, 64, 7), f"eesa_unit_{t}"),       # USE_SMOOTH_ABS 1
           (32, ("eelim_unit", t, 64, 7), f"eelim_unit_{t}"),     # EE_COST 1 + USE_LIMITS_FLAG 1
           (32, ("lim_unit", t, 64, 7), f"lim_unit_{t}"),
           (32, ("lim_trace", t, 4, 0.0, 2), f"lim_trace_{t}_N32_s4_tol0")]
    # PLANT 1-3 (oracle/ref_harness/adapt_plant.cuh): plant functions + integrator gradient on random states, traces of whole solves
    for cfg, seeds in (("p1_i3_N32_a1", (0,)), ("p1_i2_N32_a4", (1,)), ("p2_i3_N64_a8", (0, 2)), ("p2_i1_N32_a8", (1,)),
                       ("p3_i3_N64_a16", (0,)), ("p3_i2_N32_a16", (1,))):
        out.append((cfg, ("unit", t, 96, 7), f"{cfg}_unit_{t}"))
        for sd in seeds:
            out.append((cfg, ("trace", t, sd, 0.0 if sd < 2 else 0.0001, 2 if "N64" not in cfg else 1), f"{cfg}_trace_{t}_s{sd}"))
    if hw == "G":
        out += [(cfg, ("solve", "G", 0, 4, 0.0), f"{cfg}_solve_G_s0-3") for cfg in
                ("p1_i3_N32_a1", "p1_i2_N32_a4", "p2_i3_N64_a8", "p2_i1_N32_a8", "p3_i3_N64_a16", "p3_i2_N32_a16", "p3_i3_N256_a32")]
    if hw == "G":
        out += [(128, ("solve", "G", 0, 64, 0.0), "solve_G_N128_s0-63_tol0"),
                (128, ("solve", "G", 0, 64, 0.0001), "solve_G_N128_s0-63_tol1e-4"),
                (32, ("solve", "G", 0, 16, 0.0), "solve_G_N32_s0-15_tol0"),
                # warm starts of loadVarsGPU: cold solve (tol 1e-4), then (rollout, clear) = (1,0), (0,0), (1,1) from a perturbed start
                (32, ("warm", "G", 1, 0.0001, 0.0), "warm_G_N32_s1"),
                (128, ("warm", "G", 2, 0.0001, 0.0001), "warm_G_N128_s2"),
                # receding horizon: runiLQR_MPC_GPU (MPC_MODE build, gravity 0): seed, steps, knots shifted per step, iteration cap
                (32, ("mpc", 5, 8, 2, 4), "mpc_G_N32_s5"),
                (128, ("mpc", 6, 5, 3, 6), "mpc_G_N128_s6"),
                # end-effector cost: whole solves of the reference's EE_COST build
                (32, ("lim_solve", "G", 0, 8, 0.0), "lim_solve_G_N32_s0-7_tol0"),
                (32, ("eesa_solve", "G", 0, 4, 0.0), "eesa_solve_G_N32_s0-3_tol0"),
                (32, ("eelim_solve", "G", 0, 4, 0.0), "eelim_solve_G_N32_s0-3_tol0"),
                (32, ("ee_solve", "G", 0, 4, 0.0), "ee_solve_G_N32_s0-3_tol0"),
                (128, ("ee_solve", "G", 0, 2, 0.0), "ee_solve_G_N128_s0-1_tol0"),
                (32, ("ee_warm", "G", 1, 0.0001, 0.0), "ee_warm_G_N32_s1"),
                (32, ("mpc_ee", 5, 8, 2, 4), "mpc_ee_G_N32_s5"),
                (32, ("mpc_ee", 7, 6, 3, 5, "CS"), "mpc_ee_cs_G_N32_s7")]      # use_cost_shift = 1
    return out


def main():
    mode = sys.argv[1]
    if mode == "host":
        tmp = "/tmp/pddp_golden"; os.makedirs(tmp, exist_ok=True)
        for N, args, name in jobs("H"):
            b = os.path.join(tmp, name + ".bin")
            run(N, *args, b)
            to_npz(b, os.path.join(HERE, name + ".npz"))
        # trajectory hand-off message (lcmt_trajectory_f) built and encoded by the reference's generated type
        for fb in (1, 0):
            out = os.path.join(HERE, f"lcm_traj_f_N8_{'fb' if fb else 'nofb'}.bin")
            subprocess.run([os.path.join(REF, "ref_lcm_traj"), "8", str(fb), out], check=True, stdout=subprocess.DEVNULL)
        # consumer side of the hand-off: getHardwareControls on a random plan (host code of the reference)
        b = os.path.join(tmp, "hwc.bin"); subprocess.run([os.path.join(REF, "ref_hwc"), "3", b], check=True); to_npz(b, os.path.join(HERE, "hwc_N32_s3.npz"))
        d = refdump.load(os.path.join(tmp, "unit_H.bin"))
        np.savez(os.path.join(HERE, "kuka_model.npz"), I=d["I"], Tbody=d["Tbody"])
    elif mode == "gpu":
        os.makedirs(RAW, exist_ok=True)
        only = sys.argv[2] if len(sys.argv) > 2 else ""
        for N, args, name in jobs("G"):
            if only and not name.startswith(only):
                continue
            run(N, *args, os.path.join(RAW, name + ".bin"))
    elif mode == "import":
        for f in sorted(os.listdir(RAW)):
            if f.endswith(".bin"):
                to_npz(os.path.join(RAW, f), os.path.join(HERE, f[:-4] + ".npz"))
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
