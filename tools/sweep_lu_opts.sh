#!/bin/bash
# Sweep of lu.symbolic.analyze options on the bench workload (run on the GPU box); log in gpurun_out/sweep.log
mkdir -p gpurun_out; : > gpurun_out/sweep.log
while IFS= read -r opt; do
  [ -z "$opt" ] && continue
  echo "== $opt" >> gpurun_out/sweep.log
  timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --lu-opt "$opt" 2> gpurun_out/sweep.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']; r=d['roofline']['classes']
    print('it/s %.0f  nnzLU %d stored %d flops %.2e | factor %.2f ms solve %.2f ms'%(d['value'],c['nnz_LU'],c['lu_doubles_stored'],c['lu_flops_per_scenario'],r['lu_refactor']['ms_per_launch'],r['lu_solve']['ms_per_launch']))
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/sweep.err').read()[-600:])" >> gpurun_out/sweep.log
done < "${1:-tools/sweep_opts.txt}"
cat gpurun_out/sweep.log
