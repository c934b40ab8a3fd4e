#!/bin/bash
# one ncu --set full capture of a kernel: tools/gpu_prof_one.sh <kernel regex> [batch] [skip]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K="$1"; B="${2:-64}"; SKIP="${3:-1}"
PDDP_GROUPS=1 PDDP_GRAPHS=0 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/prof_${K}_b$B python tools/prof_run.py 2 $B > gpurun_out/prof_${K}_b$B.log 2>&1
tail -2 gpurun_out/prof_${K}_b$B.log; ls -la gpurun_out/prof_${K}_b$B.ncu-rep
