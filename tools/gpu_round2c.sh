#!/bin/bash
# two GPUs: full GPU suite on device 0, the 2-GPU tests, then the multi-GPU bench exactly as the driver launches it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -12
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"; tail -c 3000 gpurun_out/bench_n2.json; grep -c "NCCL INFO" gpurun_out/bench_n2.err; grep -m3 "nranks\|Init COMPLETE" gpurun_out/bench_n2.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 | cut -c1-600
