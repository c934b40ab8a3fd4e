#!/bin/bash
# warp-chain backward pass: parity (both shapes), then the kernel alone over the batch for both shapes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plugin.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -5
echo "--- block-cooperative shape: parity"; PDDP_BP_SHAPE=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "phases or whole_solve or oracle_small" 2>&1 | tail -3
echo "--- warp chains"; PDDP_BP_SHAPE=1 python tools/bp_scaling.py 1 8 64 256 1024 4096 2>&1 | tail -8; cp gpurun_out/bp_scaling.json gpurun_out/bp_scaling_warp.json
echo "--- block-cooperative"; PDDP_BP_SHAPE=2 python tools/bp_scaling.py 1 8 64 256 1024 4096 2>&1 | tail -8; cp gpurun_out/bp_scaling.json gpurun_out/bp_scaling_block.json
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_us'], d['phases_ms_per_step_unoverlapped'], d.get('strong'), d.get('reference_gpu'))"; tail -3 gpurun_out/bench.err
