#!/bin/bash
# final measurement pass: bench lines (cfg2 + reference arm, cfg3 one GPU), launch list, ncu --set full of the factor
# kernel, compute-sanitizer racecheck / memcheck of the LU kernels, smoke()
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2n_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2n_bench_cfg2.json 2> gpurun_out/r2n_bench_cfg2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2n_ref_cfg2.json 2> gpurun_out/r2n_ref_cfg2.err
python bench.py --config cfg3 --steps 2 --warmup 3 > gpurun_out/r2n_bench_cfg3.json 2> gpurun_out/r2n_bench_cfg3.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2n_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lu_factor_kernel -s 3 -c 1 -f -o gpurun_out/r2n_factor \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2n_ncu2.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_lu.py > gpurun_out/r2n_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_racecheck.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_lu.py > gpurun_out/r2n_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_memcheck.log
tail -2 gpurun_out/r2n_smoke.log; tail -4 gpurun_out/r2n_racecheck.log; tail -4 gpurun_out/r2n_memcheck.log
