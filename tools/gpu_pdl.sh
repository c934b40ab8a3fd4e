#!/bin/bash
# programmatic dependent launch on/off: parity suite, then headline rate, single-problem iteration and receding-horizon step latency
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for p in 0 1; do
  echo "== PDDP_PDL=$p"
  PDDP_PDL=$p timeout 200 python tools/quick_rate.py 5 64 2 2>&1 | tail -3
  PDDP_PDL=$p timeout 200 python tools/quick_rate.py 5 1 1 2>&1 | tail -3
  PDDP_PDL=$p timeout 200 python tools/mpc_time.py 1 5 2 2>&1 | tail -1
  PDDP_PDL=$p timeout 200 python tools/mpc_time.py 64 5 2 2>&1 | tail -1
done
