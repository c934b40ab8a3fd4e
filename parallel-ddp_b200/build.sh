#!/bin/bash
# Build libpddp.so in-tree for sm_100a (B200).
#   pddp_api.cu (host API + the Kuka kernels)  -fmad=false: every fused multiply-add on the hot path is written explicitly
#                                              (pddp_math.cuh), the compiler must not add or remove any
#   plant_tu.cu x 3 (pendulum, cart-pole, quadrotor as plug-in plants)   default contraction for the plant headers, exactly as the
#                                              reference's plant files are compiled; the solver arithmetic in them is explicit
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
NVCC="${NVCC:-nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
OBJ="$HERE/build"; mkdir -p "$OBJ"
"$NVCC" -O3 -std=c++17 $ARCH -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -v -c "$HERE/csrc/pddp_api.cu" -o "$OBJ/pddp_api.o" 2> "$HERE/build.log" &
pids=$!
for p in "1 pendulum" "2 cartpole" "3 quadrotor"; do
    set -- $p
    "$NVCC" -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC -Xptxas -v -DPDDP_PLANT_BUILTIN -DPDDP_PLANT_ID=$1 \
        -DPDDP_PLANT_HEADER="\"plants/$2.cuh\"" -DPDDP_PLANT_NAME="\"$2\"" -c "$HERE/csrc/plant_tu.cu" -o "$OBJ/plant_$1.o" 2> "$OBJ/plant_$1.log" &
    pids="$pids $!"
done
rc=0; for p in $pids; do wait $p || rc=1; done
cat "$OBJ"/plant_*.log >> "$HERE/build.log"
if [ $rc -ne 0 ]; then cat "$HERE/build.log"; exit 1; fi
"$NVCC" $ARCH -shared -Xcompiler -fPIC -o "$HERE/libpddp.so" "$OBJ/pddp_api.o" "$OBJ/plant_1.o" "$OBJ/plant_2.o" "$OBJ/plant_3.o" -ldl
grep -E "error|warning" "$HERE/build.log" | grep -v "ptxas info" | head -20 || true
