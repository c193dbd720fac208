#!/bin/bash
# last confirmation of the round: whole GPU suite + the default bench line on the final sources
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/fin2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/fin2_pytest.log
timeout 300 python bench.py > gpurun_out/fin2_bench_default.json 2> gpurun_out/fin2_bench_default.err; echo "bench rc=$?" >> gpurun_out/fin2_bench_default.err
tail -3 gpurun_out/fin2_pytest.log; tail -1 gpurun_out/fin2_bench_default.err; tail -c 1200 gpurun_out/fin2_bench_default.json
