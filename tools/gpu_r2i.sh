#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_events.py tests/test_gpu_rodas.py tests/test_gpu_adapter.py -q -p no:warnings -x > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2i_tests.log)
tail -30 gpurun_out/r2i_tests.log
