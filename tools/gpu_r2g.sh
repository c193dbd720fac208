#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -p no:warnings > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2g_tests.log)
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -8 gpurun_out/r2g_tests.log; python - <<PY
import json
d=json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], {k:round(v['ms_per_launch'],3) for k,v in d['roofline']['classes'].items()})
PY
