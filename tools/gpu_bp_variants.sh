#!/bin/bash
cd "$(dirname "$0")/.."
for v in "$@"; do echo "== $v"; PDDP_LIB=parallel-ddp_b200/libpddp_$v.so python tools/bp_scaling.py 64 1024 4096 2>&1 | tail -3; done
