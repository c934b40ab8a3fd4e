#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for g in 2 4 8; do
  PDDP_GROUPS=$g PDDP_BENCH_STRONG=0 timeout 300 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('groups', d['config']['stream_groups_per_gpu'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2))"
done
PDDP_GRAPHS=0 PDDP_BENCH_STRONG=0 timeout 300 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('no graphs: value', round(d['value']), 'e2e', round(d['e2e']['value']))"
