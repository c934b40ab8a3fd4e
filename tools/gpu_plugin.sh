#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_plugin.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_plugin.log 2>&1; echo "plugin rc=$?"; tail -25 gpurun_out/pytest_plugin.log
