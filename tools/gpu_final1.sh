#!/bin/bash
# round-2 final pass on one GPU: the whole GPU test suite, bench lines (cfg2 + reference arm, cfg3), launch list,
# ncu --set full of the factor kernel, cfg5 at the named pipe count and at the named batch, smoke()
mkdir -p gpurun_out
timeout 780 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/fin_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/fin_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/fin_bench_cfg2.json 2> gpurun_out/fin_bench_cfg2.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/fin_ref_cfg2.json 2> gpurun_out/fin_ref_cfg2.err
timeout 300 python bench.py --config cfg3 --steps 2 --warmup 3 > gpurun_out/fin_bench_cfg3.json 2> gpurun_out/fin_bench_cfg3.err
timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fin_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/fin_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lu_factor_kernel -s 3 -c 1 -f -o gpurun_out/fin_factor \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/fin_ncu2.log 2>&1
if [ -f build/models/dhs_lattice_158x158/param_and_setting.pkl ]; then
timeout 420 python tools/run_cfg5.py --nx 158 --batch 256 --steps 60 > gpurun_out/fin_cfg5_nx158.json 2> gpurun_out/fin_cfg5_nx158.err; echo "rc=$?" >> gpurun_out/fin_cfg5_nx158.err
fi
timeout 300 python tools/run_cfg5.py --nx 100 --batch 2048 --steps 60 > gpurun_out/fin_cfg5_nx100.json 2> gpurun_out/fin_cfg5_nx100.err; echo "rc=$?" >> gpurun_out/fin_cfg5_nx100.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/fin_smoke.log
tail -4 gpurun_out/fin_pytest.log; tail -2 gpurun_out/fin_smoke.log; tail -c 700 gpurun_out/fin_cfg5_nx158.json; tail -2 gpurun_out/fin_cfg5_nx158.err | cut -c1-200; tail -c 300 gpurun_out/fin_cfg5_nx100.json
