#!/bin/bash
# round-2 first GPU pass: tests, bench lines (cfg2, cfg3), reference arms, phase timing
mkdir -p gpurun_out; nproc > gpurun_out/r2a_nproc.txt
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log)
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_cfg2.json 2> gpurun_out/r2a_bench_cfg2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_ref_cfg2.json 2> gpurun_out/r2a_ref_cfg2.err
python bench.py --config cfg3 --steps 2 --warmup 3 > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err
python bench.py --config cfg3 --impl reference --steps 1 --warmup 1 > gpurun_out/r2a_ref_cfg3.json 2> gpurun_out/r2a_ref_cfg3.err
bash tools/phase_timing.sh > gpurun_out/r2a_phase.txt 2>&1
bash tools/level_timing.sh > gpurun_out/r2a_level.txt 2>&1
tail -3 gpurun_out/r2a_tests.log
