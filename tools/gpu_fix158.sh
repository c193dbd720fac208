#!/bin/bash
# after the outer-block fix: large-front tests, cfg5 at the named pipe count, fresh ncu capture of the factor kernel
# (traffic tied to the new source hash), bench lines cfg2 / cfg3 with that capture in place
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_pf.py -q -m gpu --timeout 300 -k "large_fronts or config3" > gpurun_out/fix_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/fix_pytest.log
timeout 420 python tools/run_cfg5.py --nx 158 --batch 256 --steps 60 > gpurun_out/fix_cfg5_nx158.json 2> gpurun_out/fix_cfg5_nx158.err; echo "rc=$?" >> gpurun_out/fix_cfg5_nx158.err
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lu_factor_kernel -s 3 -c 1 -f -o gpurun_out/fix_factor \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/fix_ncu2.log 2>&1
python tools/ncu_traffic.py gpurun_out/fix_factor.ncu-rep lu_refactor 32 4096 > gpurun_out/fix_traffic.log 2>&1
cp profiles/ncu_traffic.json gpurun_out/fix_ncu_traffic.json
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/fix_bench_cfg2.json 2> gpurun_out/fix_bench_cfg2.err
timeout 300 python bench.py --config cfg3 --steps 2 --warmup 3 > gpurun_out/fix_bench_cfg3.json 2> gpurun_out/fix_bench_cfg3.err
tail -3 gpurun_out/fix_pytest.log; tail -c 800 gpurun_out/fix_cfg5_nx158.json; tail -2 gpurun_out/fix_cfg5_nx158.err | cut -c1-200; tail -2 gpurun_out/fix_traffic.log | cut -c1-300
