#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -p no:warnings --timeout 300 > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2l_tests.log)
python bench.py --steps 5 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
tail -12 gpurun_out/r2l_tests.log | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r2l_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['config']['nnz_LU'], {k:round(v['ms_per_launch'],3) for k,v in d['roofline']['classes'].items()})
PY
