#!/bin/bash
# plug-in plant tests only (no -x: every tag reports)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests/test_gpu_plants.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_plants.log 2>&1; echo "plants rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/pytest_plants.log | tail -40
timeout 600 python -m pytest tests/test_gpu_shim.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_shim.log 2>&1; echo "shim rc=$?"; tail -15 gpurun_out/pytest_shim.log
cat gpurun_out/parity_report.jsonl | head -40
