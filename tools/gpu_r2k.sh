#!/bin/bash
# full GPU suite, bench line + launch list + one ncu --set full capture of the factor kernel (each ncu pass only after
# the same command exited 0 without ncu)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -p no:warnings > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2k_tests.log)
python bench.py --steps 5 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2k_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lu_factor_kernel -s 3 -c 1 -f -o gpurun_out/r2k_factor \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2k_ncu2.log 2>&1
tail -6 gpurun_out/r2k_tests.log
