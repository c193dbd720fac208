#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_vocab.py tests/test_gpu_rodas.py -q -p no:warnings --timeout 300 > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_tests.log)
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lu_backward_kernel -s 3 -c 1 -f -o gpurun_out/r2m_bwd \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2m_ncu1.log 2>&1
tail -5 gpurun_out/r2m_tests.log | cut -c1-300
