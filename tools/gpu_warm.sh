#!/bin/bash
# warm-start goldens from the reference's own GPU run, then the parity test against them (same box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/golden_raw; rm -f gpurun_out/golden_raw/*.bin
python tests/golden/make_goldens.py gpu warm 2>&1 | tail -3
python tests/golden/make_goldens.py import 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
