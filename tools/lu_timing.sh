#!/bin/bash
# Phase timing of the factor kernel: builds an instrumented libszb (-DSZB_LU_TIMING) beside the product
# library and prints cycles per group for each phase.  usage (on a GPU box): tools/lu_timing.sh [bench args]
set -e
cd "$(dirname "$0")/.."
mkdir -p build/timing
for f in szb_core szb_lu szb_rodas; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden \
       -DSZB_LU_TIMING -Iinclude -Isolverz_b200/csrc -c solverz_b200/csrc/$f.cu -o build/timing/$f.o
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/timing/libszb.so build/timing/*.o -lcudart -ldl
SZB_LIB=$PWD/build/timing/libszb.so python tools/lu_timing.py "$@"
