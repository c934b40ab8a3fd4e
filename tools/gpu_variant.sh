#!/bin/bash
# usage: tools/gpu_variant.sh lib.so ...   -- phase times of library variants on the headline configuration
cd "$(dirname "$0")/.."
for lib in "$@"; do echo "== $lib"; PDDP_LIB=$lib PDDP_GROUPS=1 python tools/prof_run.py 20 64 2>&1 | tail -1; done
