#!/bin/bash
# evidence run of the round's final code (one gpurun call): sanitizer, ncu launch lists (bench.py; one problem group kernel by kernel),
# ncu --set full of every kernel, bp traffic, bp scaling of both shapes, receding-horizon latency, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/box.txt; nproc >> gpurun_out/box.txt
bash tools/gpu_sanitize.sh > gpurun_out/sanitizer.txt 2>&1; tail -4 gpurun_out/sanitizer.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
PDDP_GRAPHS=0 PDDP_GROUPS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/prof_run.py 4 > gpurun_out/launches.log 2>&1
for k in bp_kernel sim_kernel nis_kernel sweep_kernel select_kernel; do bash tools/gpu_prof_one.sh $k 64 1 > /dev/null 2>&1; done
PDDP_BP_SHAPE=1 bash tools/gpu_prof_one.sh bp_warp_kernel 1024 1 > /dev/null 2>&1
bash tools/gpu_bp_traffic.sh 2>&1 | tail -1
PDDP_BP_SHAPE=2 python tools/bp_scaling.py 1 64 256 1024 4096 > /dev/null 2>&1; cp gpurun_out/bp_scaling.json gpurun_out/bp_scaling_block.json
PDDP_BP_SHAPE=1 python tools/bp_scaling.py 1 64 256 1024 4096 > /dev/null 2>&1; cp gpurun_out/bp_scaling.json gpurun_out/bp_scaling_warp.json
python tools/mpc_time.py 1 | tail -1; python tools/mpc_time.py 64 | tail -1
python tools/quick_rate.py 5 64 2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; head -c 700 gpurun_out/bench_n1.json; echo
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2>/dev/null; head -c 400 gpurun_out/bench_reference_arm.json; echo
ls gpurun_out | wc -l
