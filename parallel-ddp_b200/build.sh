#!/bin/bash
# Build libpddp.so in-tree for sm_100a (B200).  -fmad=false: every fused multiply-add on the hot path is written
# explicitly (pddp_math.cuh), the compiler must not add or remove any.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
NVCC="${NVCC:-nvcc}"
"$NVCC" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
    -Xcompiler -fPIC -shared -Xptxas -v \
    -o "$HERE/libpddp.so" "$HERE/csrc/pddp_api.cu" 2> "$HERE/build.log" || { cat "$HERE/build.log"; exit 1; }
grep -E "error|warning" "$HERE/build.log" | grep -v "ptxas info" | head -20 || true
