#!/bin/bash
# quick GPU pass while iterating on the LU kernels: parity of the power-flow path, one bench line, level timing
TAG="${1:-q}"
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_pf.py tests/test_gpu_api.py -x -q -k "not config3" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log)
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
bash tools/level_timing.sh > gpurun_out/${TAG}_level.txt 2>&1
bash tools/phase_timing.sh > gpurun_out/${TAG}_phase.txt 2>&1
tail -3 gpurun_out/${TAG}_tests.log; python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], {k:round(v['ms_per_launch'],3) for k,v in d['roofline']['classes'].items()})
PY
cat gpurun_out/${TAG}_level.txt
