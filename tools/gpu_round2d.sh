#!/bin/bash
# evidence run of the round's final code: sanitizer, ncu launch list of bench.py, ncu --set full of every kernel, bp traffic, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/box.txt; nproc >> gpurun_out/box.txt
bash tools/gpu_sanitize.sh > gpurun_out/sanitizer.txt 2>&1; tail -12 gpurun_out/sanitizer.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
PDDP_GRAPHS=0 PDDP_GROUPS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/prof_run.py 4 > gpurun_out/launches.log 2>&1
for k in bp_kernel sim_kernel nis_kernel sweep_kernel select_kernel; do bash tools/gpu_prof_one.sh $k 64 1 > /dev/null 2>&1; done
PDDP_BP_SHAPE=1 bash tools/gpu_prof_one.sh bp_warp_kernel 1024 1 > /dev/null 2>&1
bash tools/gpu_bp_traffic.sh 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.json
ls gpurun_out | head -80
