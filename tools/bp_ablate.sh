#!/bin/bash
# timing ablation of the backward pass: builds variants with one stage body removed (results are garbage, only time matters)
# build step (no GPU needed): tools/bp_ablate.sh build ; run step (GPU): tools/bp_ablate.sh
cd "$(dirname "$0")/.."
HERE=parallel-ddp_b200
if [ "$1" = build ]; then
  mkdir -p oracle/_ref/ablate
  for m in 0 1 2 4 8 16 32 63; do
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -DPDDP_BP_SKIP=$m -Xcompiler -fPIC -shared -o oracle/_ref/ablate/libpddp_skip$m.so $HERE/csrc/pddp_api.cu 2>/dev/null &
  done; wait; ls oracle/_ref/ablate; exit 0
fi
for m in 0 1 2 4 8 16 32 63; do echo -n "skip=$m "; PDDP_LIB=oracle/_ref/ablate/libpddp_skip$m.so python tools/bp_scaling.py 1 64 2>&1 | tr '\n' ' '; echo; done
