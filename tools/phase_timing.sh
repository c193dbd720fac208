#!/bin/bash
# Builds libszb.so with -DSZB_PHASE_TIMING into build/timing/ and prints where lu_factor_kernel spends its
# cycles (clock64, thread 0 of CTA 0).  Run on the GPU box:  bash tools/phase_timing.sh
set -e
cd "$(dirname "$0")/../solverz_b200/csrc"
mkdir -p ../../build/timing
F="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden -I../../include -I."
nvcc $F $SZB_EXTRA -DSZB_PHASE_TIMING -c szb_lu.cu -o ../../build/timing/szb_lu.o
cp libszb.so ../../build/timing/libszb.keep.so
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libszb.so szb_core.o ../../build/timing/szb_lu.o szb_rodas.o -lcudart -ldl
cd ../..
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e "$@" 2>&1 | grep -A13 "lu_factor phases" | tail -14
cp build/timing/libszb.keep.so solverz_b200/csrc/libszb.so
