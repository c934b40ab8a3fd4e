#!/bin/bash
# usage: tools/gpu_profile.sh [kernels...]   (default: all five)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KERNELS="${@:-bp_kernel sim_kernel nis_kernel sweep_kernel select_kernel}"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/prof_run.py 4 > gpurun_out/launches.log 2>&1
for k in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k python tools/prof_run.py 2 > gpurun_out/prof_$k.log 2>&1
done
ls gpurun_out/
