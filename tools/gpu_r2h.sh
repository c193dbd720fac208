#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_pivot.py -q -p no:warnings > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2h_tests.log)
timeout 1200 python tools/run_cfg4.py --nx 45 --batch 1024 --tend 2.0 > gpurun_out/r2h_cfg4_full.json 2> gpurun_out/r2h_cfg4_full.err
echo "cfg4 rc=$?" >> gpurun_out/r2h_cfg4_full.err
tail -4 gpurun_out/r2h_tests.log; tail -c 1500 gpurun_out/r2h_cfg4_full.json; tail -3 gpurun_out/r2h_cfg4_full.err
