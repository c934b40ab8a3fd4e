#!/bin/bash
# parity tests + phase times of the headline configuration (20 iterations, one problem group)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
PDDP_GROUPS=1 python tools/prof_run.py 20 64 2>&1 | tail -1
