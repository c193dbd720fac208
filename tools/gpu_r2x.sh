#!/bin/bash
mkdir -p gpurun_out
timeout 420 python tools/run_cfg5.py --nx 100 --batch 2048 --steps 60 > gpurun_out/r2x_cfg5_nx100.json 2> gpurun_out/r2x_cfg5_nx100.err; echo "rc=$?" >> gpurun_out/r2x_cfg5_nx100.err
tail -c 1700 gpurun_out/r2x_cfg5_nx100.json; tail -3 gpurun_out/r2x_cfg5_nx100.err | cut -c1-300
