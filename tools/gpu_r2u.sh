#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/run_cfg5.py --nx 158 --batch 256 --steps 60 > gpurun_out/r2u_cfg5_nx158.json 2> gpurun_out/r2u_cfg5_nx158.err; echo "rc=$?" >> gpurun_out/r2u_cfg5_nx158.err
timeout 600 python tools/run_cfg5.py --nx 100 --batch 2048 --steps 60 > gpurun_out/r2u_cfg5_nx100.json 2> gpurun_out/r2u_cfg5_nx100.err; echo "rc=$?" >> gpurun_out/r2u_cfg5_nx100.err
tail -c 1500 gpurun_out/r2u_cfg5_nx158.json; tail -3 gpurun_out/r2u_cfg5_nx158.err | cut -c1-300; tail -c 1500 gpurun_out/r2u_cfg5_nx100.json; tail -3 gpurun_out/r2u_cfg5_nx100.err | cut -c1-300
