#!/bin/bash
# round 2, first GPU call: reference GPU dumps of the plug-in plants, the plug-in plant tests, then the whole GPU suite and a bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 600 python tests/golden/make_goldens.py gpu p > gpurun_out/make_goldens_plants.log 2>&1; echo "goldens rc=$?"; tail -3 gpurun_out/make_goldens_plants.log
timeout 900 python -m pytest tests/test_gpu_plants.py -m gpu -q -x --no-header > gpurun_out/pytest_plants.log 2>&1; echo "plants rc=$?"; tail -25 gpurun_out/pytest_plants.log
timeout 1200 python -m pytest tests -m gpu -q --no-header --deselect tests/test_gpu_plants.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
