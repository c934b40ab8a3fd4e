#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tools/bp_trace.py 64 2>&1 | tail -3
python tools/bp_scaling.py 2>&1 | tail -8
