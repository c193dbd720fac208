#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_pf.py -q -p no:warnings --timeout 400 -k "large_fronts or config3 or lu_refactor" > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2p_tests.log)
python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_bench_cfg3.json 2> gpurun_out/r2p_bench_cfg3.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_bench_cfg2.json 2> gpurun_out/r2p_bench_cfg2.err
tail -4 gpurun_out/r2p_tests.log | cut -c1-200
python - <<PY
import json
for f in ('r2p_bench_cfg3','r2p_bench_cfg2'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, round(d['value'],1), round(d['e2e']['value'],1), d['config']['converged'], {k:round(v['ms_per_launch'],3) for k,v in d['roofline']['classes'].items()})
PY
