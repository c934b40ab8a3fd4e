#!/bin/bash
# round 2: whole GPU suite (Kuka + plug-in plants + shim + plug-in library), bench line, receding-horizon latency with / without graphs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -20
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for b in 1 64; do python tools/mpc_time.py $b 5 2; PDDP_GRAPHS=0 python tools/mpc_time.py $b 5 2; done
