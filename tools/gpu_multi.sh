#!/bin/bash
# usage: tools/gpu_multi.sh N   -- bench on N GPUs (torchrun), plus the reference arm
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cat gpurun_out/bench_n$N.json | cut -c1-600; tail -3 gpurun_out/bench_n$N.err
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
timeout 300 python -m pytest tests/test_gpu_shim.py -m gpu -q 2>&1 | tail -3
