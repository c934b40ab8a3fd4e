#!/usr/bin/env python3
"""bench.py -- iLQR iterations/sec of the batched hot path on N B200s (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: pddp_solve of 64 Kuka problems (N=128 knots, 16 alphas, M=4,
TOL_COST=0 so every problem runs MAX_ITER=100 iterations) = 6400 iLQR iterations per GPU per step.

  value  : whole-job iterations/s with the inputs already resident in HBM (pddp_solve_device), device time from CUDA
           events recorded by the library on ITS launch stream, max over ranks
  e2e    : the same metric through the reference-facing C-ABI call pddp_solve with HOST buffers (H2D of x0,u0,xGoal and
           D2H of x,u,Jout,alphaOut,iters inside the timed region), wall clock around the call, max over ranks
  roofline: the backward-pass kernel (the metric's kernel): algorithmic bytes (656 452 B per problem per
           backward pass at N=128,M=4) x problems per launch / its mean launch duration, against MEASURED_PEAKS.json
  cpu_baseline: the UNMODIFIED reference's CPU path (oracle/_ref, runiLQR_CPU2 = parallel line search) on a bounded sample

`--impl reference` times that CPU path alone (rank 0 only).  Multi-GPU: problems are independent, so ranks shard the
global batch (64 per rank, weak scaling) with no data-path collective; iteration counts are all-gathered at the end."""
import argparse
import ctypes as C
import importlib
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_KNOTS, N_ALPHA, M_BLOCKS, BATCH_PER_GPU, MAX_ITER = 128, 16, 4, int(os.environ.get("PDDP_BENCH_BATCH", "64")), 100   # 64 = the headline; the env is for the strong-scaling data point (512 on one GPU)
BP_BYTES_PER_PROBLEM = 4 * ((N_KNOTS - 1) * 1281 + 4 * 238 + 3 * 14 + 2 * 210 + 3 * 4)   # = 656452
STRONG_BATCH = 512          # BASELINE configs[3]
METRIC = "ilqr_iterations_per_sec"
UNIT = "iterations/s"
CONFIG = {"workload": f"configs[2]: Kuka iiwa14 N=128 knots, alpha=16, M=4, batch={BATCH_PER_GPU} per GPU, TOL_COST=0 (100 iterations per problem)",
          "plant": "kuka_iiwa14", "knots": N_KNOTS, "n_alpha": N_ALPHA, "m_blocks": M_BLOCKS, "batch_per_gpu": BATCH_PER_GPU,
          "max_iter": MAX_ITER, "l2": f"a 256 MiB buffer is rewritten between timed steps (working set {1.08*BATCH_PER_GPU:.0f} MB)",
          # every iteration does the reference's full work (rejected line searches included) unless PDDP_SKIP_UNCHANGED=1 is set
          "skip_unchanged_gradient_refresh": bool(int(os.environ.get("PDDP_SKIP_UNCHANGED", "0") or 0))}
# secondary data point (not the headline): the same shape under the reference's end-effector cost (EE_COST 1, default weights, the
# goal pose of WAFR_iLQR_examples.cu:37-43); the reference arm and the CPU baseline stay on the joint-space cost
BENCH_EE = bool(int(os.environ.get("PDDP_BENCH_EE", "0") or 0))
if BENCH_EE:
    CONFIG["workload"] += " -- end-effector cost (EE_COST 1)"; CONFIG["cost"] = "end_effector"


def kernel_source_hash():
    import hashlib
    h = hashlib.sha1()
    for f in ("kernels.cuh", "dev_state.cuh", "pddp_math.cuh"):
        h.update(open(os.path.join(ROOT, "parallel-ddp_b200", "csrc", f), "rb").read())
    return h.hexdigest()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


_REF_FAIL = ""


def ref_cpu_run(nseeds, mode="P", seed0=0):
    """Unmodified reference CPU path on `nseeds` benchmark problems; returns (iters_per_sec, total_iters, seconds, cores) or None."""
    exe = os.path.join(ROOT, "oracle", "_ref", f"ref_driver_N{N_KNOTS}")
    if not os.path.exists(exe):
        return None
    r = subprocess.run([exe, "time", mode, str(seed0), str(nseeds), "0.0"], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    m = re.search(r"REFSUMMARY (\{.*\})", r.stderr)
    if not m:
        global _REF_FAIL
        _REF_FAIL = f"rc={r.returncode}: " + " | ".join(r.stderr.strip().splitlines()[-2:])[:200]
        return None
    s = json.loads(m.group(1))
    return s["total_iters"] / (s["sum_solve_ms"] / 1000.0), s["total_iters"], s["sum_solve_ms"] / 1000.0, s["cores"]


def oracle_port_run(nprob):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    pddp = importlib.import_module("parallel-ddp_b200")
    x0, u0, xg = pddp.make_inputs_kuka(N_KNOTS, nprob, 0)
    L = ol.lib(True); cfg = ol.kuka_cfg(N_KNOTS, fma=True); cp = C.byref(cfg)
    t0 = time.time(); tot = 0
    for b in range(nprob):
        ox = np.zeros((N_KNOTS, 14), np.float32); ou = np.zeros((N_KNOTS, 7), np.float32)
        oJ = np.zeros(MAX_ITER + 1, np.float32); oa = np.zeros(MAX_ITER + 1, np.int32)
        tot += L.orc_solve(cp, ol.fptr(x0[b]), ol.fptr(u0[b]), ol.fptr(xg[b]), ol.fptr(ox), ol.fptr(ou), ol.fptr(oJ), ol.iptr(oa))
    dt = time.time() - t0
    return tot / dt, tot, dt, 1


def ref_gpu_run(nseeds=8):
    """The UNMODIFIED reference's own GPU build (runiLQR_GPU, compiled for sm_100 by oracle/Makefile) on the first `nseeds` benchmark
    problems, one after the other as the reference solves them: north_star's ">= 10x the reference GPU build" denominator."""
    exe = os.path.join(ROOT, "oracle", "_ref", f"ref_driver_N{N_KNOTS}")
    if not os.path.exists(exe):
        return None
    best = None
    for _ in range(2):                                   # the first call pays context creation inside its first solve
        r = subprocess.run([exe, "time", "G", "0", str(nseeds), "0.0"], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        m = re.search(r"REFSUMMARY (\{.*\})", r.stderr)
        if not m:
            return None
        d = json.loads(m.group(1)); v = d["total_iters"] / (d["sum_solve_ms"] / 1000.0)
        best = max(best or 0.0, v)
    return {"value": best, "unit": UNIT, "kind": "reference GPU build (oracle/_ref/ref_driver_N128, mode G, unmodified reference kernels on this GPU)",
            "sample": f"{nseeds} benchmark problems (seeds 0..{nseeds-1}) x 100 iterations, solved one after the other (the reference has no batch), best of 2 runs"}


def cpu_baseline(nseeds=16):      # about 11 s of host work on the GPU box (16 threads)
    r = ref_cpu_run(nseeds)
    if r:
        return {"value": r[0], "unit": UNIT, "cores": r[3], "kind": "reference",
                "sample": f"{nseeds} of the 64 benchmark problems (seeds 0..{nseeds-1}) x 100 iterations, reference runiLQR_CPU2 (std::thread parallel line search), {r[2]:.1f} s"}
    r = oracle_port_run(2)
    return {"value": r[0], "unit": UNIT, "cores": 1, "kind": "port", "sample": f"2 benchmark problems x 100 iterations, single-threaded oracle port, {r[2]:.1f} s"
            + (f" (the reference binary did not run on this host: {_REF_FAIL})" if _REF_FAIL else "")}


def run_reference(args, rank):
    if rank != 0:
        return
    per_step = 2
    for _ in range(min(args.warmup, 1)):
        ref_cpu_run(1)
    tot_it, tot_s, cores, kind = 0, 0.0, 1, "reference"
    for s in range(args.steps):
        r = ref_cpu_run(per_step, seed0=(s * per_step) % 64)
        if r is None:
            r = oracle_port_run(1); kind = "port"
        tot_it += r[1]; tot_s += r[2]; cores = r[3]
    v = tot_it / tot_s
    sample = f"{per_step} benchmark problems x 100 iterations per step, {args.steps} steps"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": 1000.0 * tot_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": CONFIG,
                      "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG is left as the caller set it (the driver reads the rank count from NCCL's INFO lines).  NCCL writes those to STDOUT
        # unless told otherwise, and stdout carries exactly one JSON line here: without a NCCL_DEBUG_FILE they go to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pddp = importlib.import_module("parallel-ddp_b200")
    B, N, L1 = BATCH_PER_GPU, N_KNOTS, MAX_ITER + 1
    cfg = pddp.default_config_kuka(N, B, device=local_rank, tol_cost=0.0, ee_cost=1 if BENCH_EE else 0)
    solver = pddp.Solver(cfg)
    # this rank's shard of the global batch (seed = problem index): problems lo .. hi-1
    sharding = importlib.import_module("parallel-ddp_b200.sharding")
    lo, hi = sharding.shard_range(rank, world, B * world)
    x0, u0, xg = pddp.make_inputs_kuka(N, hi - lo, seed0=lo)
    if BENCH_EE:
        xg[:] = 0; xg[:, :6] = (0.3638, 0.0, 1.0628, 0.5 * 3.14159, 0.0, 0.5 * 3.14159)
    dev = torch.device("cuda", local_rank)
    d_x0 = torch.from_numpy(x0).to(dev); d_u0 = torch.from_numpy(u0).to(dev); d_xg = torch.from_numpy(xg).to(dev)
    d_x = torch.empty_like(d_x0); d_u = torch.empty_like(d_u0)
    d_J = torch.empty((B, L1), dtype=torch.float32, device=dev); d_a = torch.empty((B, L1), dtype=torch.int32, device=dev); d_it = torch.empty(B, dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    times = np.zeros(6, np.float64)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        flush.fill_(1); torch.cuda.synchronize()
        solver.solve_device(d_x0.data_ptr(), d_u0.data_ptr(), d_xg.data_ptr(), d_x.data_ptr(), d_u.data_ptr(), d_J.data_ptr(), d_a.data_ptr(), d_it.data_ptr(), 1, times)
        return times.copy()

    groups = solver.set_groups(int(os.environ.get("PDDP_GROUPS", "2")))
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    t_wall0 = time.time(); dev_ms = 0.0
    for _ in range(args.steps):
        t = step_device(); dev_ms += t[0]
    barrier()
    wall_s = time.time() - t_wall0
    launches = solver.launch_count() * args.steps; graph_launches = solver.graph_launch_count() * args.steps
    # per-phase device times (and the backward-pass launch duration for the roofline) need the phases un-overlapped:
    # a second, shorter pass with one problem group, every phase bracketed by CUDA events on the launch stream
    solver.set_groups(1)
    step_device()
    phase = np.zeros(6); nphase = max(2, args.steps // 2)
    for _ in range(nphase):
        phase += step_device()
    phase *= args.steps / nphase          # scaled to `steps` steps so that the keys below stay per-step after division
    solver.set_groups(groups)
    iters_rank = int(d_it.sum().item()) * args.steps
    # end-to-end through the reference-facing call with host buffers
    for _ in range(min(args.warmup, 2)):
        solver.runiLQR_GPU(x0, u0, xg)
    barrier()
    t0 = time.time(); e2e_iters = 0
    for _ in range(args.steps):
        flush.fill_(1); torch.cuda.synchronize()
        o = solver.runiLQR_GPU(x0, u0, xg); e2e_iters += int(o["iters"].sum())
    barrier()
    e2e_s = time.time() - t0
    clocks = sampler.stop()
    # informational, rank 0 at N=1 only and NOT the headline: the opt-in pddp_set_skip_unchanged -- after a rejected line search the
    # trajectory, hence AB / H / g, are unchanged, and the gradient refresh returns at once instead of recomputing identical values as
    # the reference does (bit-identical results; 47 % of this workload's iterations reject)
    optin = None
    if world == 1 and not BENCH_EE and not CONFIG["skip_unchanged_gradient_refresh"]:
        solver.set_skip_unchanged(1)
        step_device(); ms3 = 0.0; n3 = max(2, args.steps // 2)
        for _ in range(n3):
            ms3 += step_device()[0]
        optin = {"value": int(d_it.sum().item()) * n3 / (ms3 / 1000.0), "unit": UNIT, "steps": n3,
                 "note": "pddp_set_skip_unchanged(1): no gradient refresh after a rejected line search; not the default, not the headline"}
        solver.set_skip_unchanged(0)
    # strong scaling (BASELINE configs[3]): a FIXED global batch of 512 problems split over the ranks, same timing rules; on one GPU
    # this is the 512-problem single-GPU figure the N-GPU values are divided by
    strong = None
    if not BENCH_EE and os.environ.get("PDDP_BENCH_STRONG", "1") != "0" and STRONG_BATCH % world == 0:
        sb = STRONG_BATCH // world
        s2 = pddp.Solver(pddp.default_config_kuka(N, sb, device=local_rank, tol_cost=0.0))
        lo2, hi2 = sharding.shard_range(rank, world, STRONG_BATCH)
        a0, b0_, g0 = pddp.make_inputs_kuka(N, hi2 - lo2, seed0=lo2)
        e_x0 = torch.from_numpy(a0).to(dev); e_u0 = torch.from_numpy(b0_).to(dev); e_xg = torch.from_numpy(g0).to(dev)
        e_x = torch.empty_like(e_x0); e_u = torch.empty_like(e_u0)
        e_J = torch.empty((sb, L1), dtype=torch.float32, device=dev); e_a = torch.empty((sb, L1), dtype=torch.int32, device=dev); e_it = torch.empty(sb, dtype=torch.int32, device=dev)
        t2 = np.zeros(6, np.float64); nst = max(2, args.steps // 2)

        def step_strong():
            flush.fill_(1); torch.cuda.synchronize()
            s2.solve_device(e_x0.data_ptr(), e_u0.data_ptr(), e_xg.data_ptr(), e_x.data_ptr(), e_u.data_ptr(), e_J.data_ptr(), e_a.data_ptr(), e_it.data_ptr(), 1, t2)
            return t2[0]
        for _ in range(3):
            step_strong()
        barrier(); ms2 = 0.0
        for _ in range(nst):
            ms2 += step_strong()
        barrier()
        st2 = torch.tensor([ms2, float(int(e_it.sum().item()) * nst)], dtype=torch.float64, device=dev)
        if world > 1:
            m2 = st2.clone(); dist.all_reduce(m2, op=dist.ReduceOp.MAX); q2 = st2.clone(); dist.all_reduce(q2, op=dist.ReduceOp.SUM)
            ms2, it2 = m2[0].item(), q2[1].item()
        else:
            it2 = st2[1].item()
        strong = {"global_batch": STRONG_BATCH, "batch_per_gpu": sb, "n_gpus": world, "value": it2 / (ms2 / 1000.0), "unit": UNIT, "steps": nst,
                  "ms_per_step": ms2 / nst, "scaling": "strong", "note": "the 1-GPU run of this leg is the denominator of the N-GPU speed-up"}
        s2.freeMemory_GPU(); del e_x0, e_u0, e_xg, e_x, e_u, e_J, e_a, e_it
    # the other multi-GPU mode (north_star): ONE problem, its line search's step sizes sharded over the ranks, one NCCL exchange at selection
    # (libpddp's own communicator: pddp_alpha_shard_init).  Reported as latency per iteration next to the collective's share.
    ashard = None
    if world > 1 and not BENCH_EE and N_ALPHA % world == 0 and os.environ.get("PDDP_BENCH_ALPHA_SHARD", "1") != "0":
        uid = [pddp.alpha_shard_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, device=dev)
        s3 = pddp.Solver(pddp.default_config_kuka(N, 1, device=local_rank, tol_cost=0.0, max_iter=50))
        s3.alpha_shard_init(rank, world, uid[0])
        a1, b1, g1 = pddp.make_inputs_kuka(N, 1, seed0=0)
        s3.runiLQR_GPU(a1, b1, g1); barrier()
        t0 = time.time(); reps = 5
        for _ in range(reps):
            o3 = s3.runiLQR_GPU(a1, b1, g1)
        barrier()
        t3 = torch.tensor([(time.time() - t0) / reps, s3.alpha_shard_stats()["exchange_us_per_iteration"]], dtype=torch.float64, device=dev)
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        ashard = {"workload": "1 problem, Kuka N=128, alpha=16 split over the ranks, 50 iterations, host buffers", "n_gpus": world, "step_sizes_per_gpu": N_ALPHA // world,
                  "ms_per_iteration": 1000.0 * t3[0].item() / 50, "nccl_exchange_us_per_iteration": t3[1].item(),
                  "collectives_per_iteration": "1 ncclAllGather of 2*alpha floats (selection) + 1 ncclAllReduce of the accepted candidate (4 N (2n+m) B)",
                  "final_cost": float(o3["Jout"][0, 50])}
        s3.freeMemory_GPU()
    h2d = x0.nbytes + u0.nbytes + xg.nbytes
    d2h = o["x"].nbytes + o["u"].nbytes + o["Jout"].nbytes + o["alphaOut"].nbytes + o["iters"].nbytes
    stats = torch.tensor([dev_ms, e2e_s, float(iters_rank), float(e2e_iters), phase[3]], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s, bp_ms = mx[0].item(), mx[1].item(), mx[4].item()
        iters_all, e2e_iters_all = sm[2].item(), sm[3].item()
        all_iters = sharding.gather_counters(d_it)      # result hand-off (per-problem iteration counters) over NCCL / NVLink
        assert all_iters.numel() == B * world
    else:
        iters_all, e2e_iters_all, bp_ms = float(iters_rank), float(e2e_iters), phase[3]
    if rank == 0:
        value = iters_all / (dev_ms / 1000.0)
        bp_launches = MAX_ITER * args.steps
        bp_avg_s = (bp_ms / 1000.0) / bp_launches
        peak, peak_src = peaks()
        achieved = B * BP_BYTES_PER_PROBLEM / bp_avg_s / 1e9
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": dict(CONFIG, global_batch=B * world, parallelism=f"problem-sharded x{world}", stream_groups_per_gpu=groups),
               "e2e": {"value": e2e_iters_all / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
               "gpu_launches": launches, "graph_launches": graph_launches, "clocks": clocks,
               "roofline": {"kernel": "bp_kernel<14,7>", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": None, "peak_source": peak_src, "avg_launch_us": bp_avg_s * 1e6, "problems_per_launch": B,
                            "algorithmic_bytes_per_problem": BP_BYTES_PER_PROBLEM},
               "phases_ms_per_step_unoverlapped": {k: float(v) / args.steps for k, v in zip(("total", "sim+select", "sweep", "bp", "nis", "init+store"), phase)},
               "wall_s_device_leg": wall_s}
        # the two kernels that take most of an iteration are FP32-issue / latency bound: algorithmic FLOP of the
        # reference's formulation (FMA = 2) per unit, counted with an instrumented build of the oracle (every FMA/MUL/ADD/SUB of
        # oracle/pddp_oracle.c): forward dynamics + Euler step 12 628, + control update and cost ~= 12.9 kFLOP per (candidate,
        # knot); analytic integrator gradient 184 059 ~= 184 kFLOP per knot --
        # against the nominal FP32 peak 148 SM x 128 FMA/clk x 2 x SM clock
        ph = out["phases_ms_per_step_unoverlapped"]; sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        sim_flop = B * N_ALPHA * (N_KNOTS - 1) * 12.9e3 * MAX_ITER; nis_flop = B * N_KNOTS * 184e3 * MAX_ITER
        out["other_kernels"] = [
            {"kernel": "sim_kernel (+select)", "bound": "shared-memory pipe / fp32 issue (ncu: 67 % / 59 % of peak)", "achieved": sim_flop / (ph["sim+select"] * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
             "frac": sim_flop / (ph["sim+select"] * 1e-3) / 1e12 / fp32_peak, "peak_source": "nominal FP32 (no tensor cores: bit-exact fp32 chains)"},
            {"kernel": "nis_kernel", "bound": "fp32 issue / latency (16 warps per SM)", "achieved": nis_flop / (ph["nis"] * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
             "frac": nis_flop / (ph["nis"] * 1e-3) / 1e12 / fp32_peak, "peak_source": "nominal FP32 (no tensor cores: bit-exact fp32 chains)"}]
        # DRAM traffic of the backward pass: one ncu --set full capture per kernel version (tools/gpu_bp_traffic.sh), not something a bench
        # run can measure itself; the capture records the hash of the kernel source it was taken from, and a capture of an older
        # source is reported as stale instead of passing for the current kernel
        ncu = os.path.join(ROOT, "profiles", "bp_traffic.json")
        if os.path.exists(ncu):
            tr = json.load(open(ncu)); cur = kernel_source_hash()
            out["roofline"]["traffic"] = tr.get("dram_bytes_per_launch")
            out["roofline"]["traffic_source"] = {"file": "profiles/bp_traffic.json", "captured_from_source_sha1": tr.get("kernel_source_sha1"),
                                                 "current_source_sha1": cur, "stale": tr.get("kernel_source_sha1") != cur}
        # the same kernel where it saturates: the backward pass alone over 2048 problems (8192 chains: the warp-chain shape), timed by the
        # library's CUDA events through the phase-level entry point
        if world == 1 and os.environ.get("PDDP_BENCH_LARGE", "1") != "0":
            LB = 2048
            s4 = pddp.Solver(pddp.default_config_kuka(N, LB, device=local_rank, tol_cost=0.0, max_iter=3))
            a4, b4, g4 = pddp.make_inputs_kuka(N, 64, seed0=0)
            s4.load_init(np.tile(a4, (LB // 64, 1, 1)), np.tile(b4, (LB // 64, 1, 1)), np.tile(g4, (LB // 64, 1)))
            ts = []
            for it in range(7):
                flush.fill_(1); torch.cuda.synchronize()
                s4.backwardPassGPU(); ts.append(s4.last_phase_ms()[0])
                if it < 2:
                    s4.forwardSweep(); s4.forwardSimGPU(); s4.nextIterationSetupGPU()
            t4 = float(np.median(ts[2:])) * 1e-3
            out["roofline_large_batch"] = {"kernel": "bp_warp_kernel (warp chains, chosen by launch size)", "bound": "hbm", "problems_per_launch": LB,
                                           "achieved": LB * BP_BYTES_PER_PROBLEM / t4 / 1e9, "peak": peak, "unit": "GB/s",
                                           "frac": LB * BP_BYTES_PER_PROBLEM / t4 / 1e9 / peak, "avg_launch_us": t4 * 1e6, "peak_source": peak_src}
            s4.freeMemory_GPU()
        if strong is not None:
            out["strong"] = strong
        if optin is not None:
            out["optin_skip_unchanged"] = optin
        if ashard is not None:
            out["alpha_sharded"] = ashard
        if world == 1:
            out["cpu_baseline"] = cpu_baseline()
            rg = ref_gpu_run()
            if rg:
                rg["speedup_value"] = value / rg["value"]; rg["speedup_e2e"] = out["e2e"]["value"] / rg["value"]
                out["reference_gpu"] = rg
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
