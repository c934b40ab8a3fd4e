#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "phases or whole_solve or oracle or mpc or config3" 2>&1 | tail -3
echo "--- warp chains"; PDDP_BP_SHAPE=1 python tools/bp_scaling.py 64 256 1024 4096 2>&1 | tail -4
