#!/bin/bash
# full GPU test suite + one ncu --set full capture of the factor kernel (after the bench command exited 0 without ncu)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q -p no:warnings > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log)
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err && \
ncu --set full --clock-control none --import-source on -k regex:lu_factor_kernel -s 3 -c 1 -f -o gpurun_out/r2f_factor \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2f_ncu.log 2>&1
tail -5 gpurun_out/r2f_tests.log; ls -la gpurun_out/r2f_factor.ncu-rep
