#!/bin/bash
# end-effector cost: reference GPU dumps (ee_unit_G, ee_solve_G) + CUDA path / oracle comparison
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tests/golden/make_goldens.py gpu ee_ 2>&1 | tail -3
timeout 300 python tools/ee_check.py 2>&1 | tail -30
