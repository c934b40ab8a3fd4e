#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for g in 1 2 4; do PDDP_GROUPS=$g python bench.py --steps 5 --warmup 3 > gpurun_out/bench_g$g.json 2>gpurun_out/bench_g$g.err; python -c "
import json; d=json.load(open('gpurun_out/bench_g$g.json')); print('groups', $g, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],2), 'bp_us', round(d['roofline']['avg_launch_us'],1))"; done
cp gpurun_out/bench_g2.json gpurun_out/bench.json
