#!/bin/bash
# experiment: 512-thread CTAs (128 registers per thread) for the factor kernel
set -e
cd solverz_b200/csrc
F="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden -I../../include -I."
cp libszb.so /tmp/keep.so
nvcc $F -DSZB_FACTOR_THREADS=512 -Xptxas -v -c szb_lu.cu -o /tmp/szb_lu512.o 2>&1 | grep -A2 lu_factor | grep -i "spill\|registers"
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libszb.so szb_core.o /tmp/szb_lu512.o szb_rodas.o -lcudart -ldl
cd ../..
for opt in '{"n_warps":16}' '{"n_warps":16,"n_groups":2}' '{"n_warps":16,"n_groups":4,"warp_front_max":32}'; do
echo "== $opt"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --lu-opt "$opt" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']; r=d['roofline']['classes']
print('it/s %.0f | factor %.2f ms solve %.2f ms'%(d['value'],r['lu_refactor']['ms_per_launch'],r['lu_solve']['ms_per_launch']))"
done
cp /tmp/keep.so solverz_b200/csrc/libszb.so
