#!/bin/bash
# DRAM traffic of one refactor launch at cfg3 (fronts in global memory), then the cfg3 bench line with it in place
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 200 python bench.py --config cfg3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1 && \
timeout 400 ncu --metrics $M --clock-control none -k regex:lu_factor_kernel -s 2 -c 1 -f -o gpurun_out/cfg3_factor \
    python bench.py --config cfg3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/cfg3_ncu.log 2>&1
python tools/ncu_traffic.py gpurun_out/cfg3_factor.ncu-rep lu_refactor 100 2048 "ncu --metrics $M --clock-control none, one launch" > gpurun_out/cfg3_traffic.log 2>&1
cp profiles/ncu_traffic.json gpurun_out/cfg3_ncu_traffic.json
timeout 300 python bench.py --config cfg3 --steps 2 --warmup 3 > gpurun_out/cfg3_bench.json 2> gpurun_out/cfg3_bench.err
tail -2 gpurun_out/cfg3_traffic.log | cut -c1-400; tail -c 300 gpurun_out/cfg3_bench.json
