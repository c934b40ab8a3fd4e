#!/bin/bash
# receding-horizon goldens from the reference's own GPU run (ref_mpc_N*), then the parity test against them (same box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/golden_raw; rm -f gpurun_out/golden_raw/*.bin
if [ ! -f tests/golden/mpc_G_N32_s5.npz ]; then
  python tests/golden/make_goldens.py gpu mpc 2>&1 | tail -3
  python tests/golden/make_goldens.py import 2>&1 | tail -3
fi
timeout 150 python -m pytest tests -m gpu -q -x -k mpc 2>&1 | tail -25
