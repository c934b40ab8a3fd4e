#!/bin/bash
# first GPU call: reference GPU goldens + reference timings + sanitizer + parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> gpurun_out/gpu.txt
echo "== goldens" ; ( time python tests/golden/make_goldens.py gpu ) > gpurun_out/make_goldens.log 2>&1; tail -3 gpurun_out/make_goldens.log
echo "== reference timings"
( oracle/_ref/ref_driver_N128 time G 0 64 0.0 > /dev/null ) 2> gpurun_out/ref_time_G.log; grep REFSUMMARY gpurun_out/ref_time_G.log
( oracle/_ref/ref_driver_N128 time C 0 4 0.0 > /dev/null ) 2> gpurun_out/ref_time_C.log; grep REFSUMMARY gpurun_out/ref_time_C.log
( oracle/_ref/ref_driver_N128 time P 0 4 0.0 > /dev/null ) 2> gpurun_out/ref_time_P.log; grep REFSUMMARY gpurun_out/ref_time_P.log
echo "== sanitizer"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/smoke_small.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
echo "== pytest"
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
