#!/bin/bash
# two GPUs: the sharded line search against the single-GPU solve, then single-problem latency with 1 and 2 GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_alpha_shard.py -m gpu -q -s --no-header -p no:cacheprovider 2>&1 | tail -15
timeout 600 python tools/alpha_shard_latency.py 2>&1 | tail -6
