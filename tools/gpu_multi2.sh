#!/bin/bash
# multi-GPU bench lines as the driver launches them: tools/gpu_multi2.sh <N>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/nccl_n${N}_%p.log timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench$N rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, 'e2e', d['e2e']['value']); print(d.get('strong')); print(d.get('alpha_sharded'))"
grep -h "nranks" gpurun_out/nccl_n${N}_*.log | head -3; ls gpurun_out/nccl_n${N}_*.log | wc -l; rm -f gpurun_out/nccl_n${N}_*.log
PDDP_LAT_BATCH=1 timeout 600 python tools/alpha_shard_latency.py 2>&1 | tail -1
