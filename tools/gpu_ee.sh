#!/bin/bash
# end-effector cost: reference GPU dumps (ee_unit_G, ee_solve_G, ee_warm_G) for tests/golden, parity tests, phase times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tests/golden/make_goldens.py gpu ee_ 2>&1 | tail -5
timeout 300 python tests/golden/make_goldens.py gpu mpc_ee 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
PDDP_GROUPS=1 python tools/prof_run.py 20 64 2>&1 | tail -1
PDDP_EE=1 PDDP_GROUPS=1 python tools/prof_run.py 20 64 2>&1 | tail -1
