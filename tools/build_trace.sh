#!/bin/bash
# trace build of the library: the backward pass records clock64() per stage into DevState.dbg (tools/bp_trace.py), the forward
# dynamics of the simulation per phase into pddp_simtrace (tools/sim_trace.py); needs build.sh to have run (plant objects)
HERE="$(cd "$(dirname "$0")/../parallel-ddp_b200" && pwd)"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -DPDDP_BP_TRACE -DPDDP_SIM_TRACE -DPDDP_TRACE_THREAD=${TRACE_THREAD:-0} -Xcompiler -fPIC -shared \
    -o "$HERE/libpddp_trace.so" "$HERE/csrc/pddp_api.cu" "$HERE/build/plant_1.o" "$HERE/build/plant_2.o" "$HERE/build/plant_3.o" -ldl 2>/dev/null
