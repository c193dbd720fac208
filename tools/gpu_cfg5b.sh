#!/bin/bash
mkdir -p gpurun_out
timeout 280 python tools/run_cfg5.py --nx 158 --batch 256 --steps 60 --tol 1e-6 --max-it 12 > gpurun_out/fix_cfg5_nx158.json 2> gpurun_out/fix_cfg5_nx158.err; echo "rc=$?" >> gpurun_out/fix_cfg5_nx158.err
tail -c 1500 gpurun_out/fix_cfg5_nx158.json; tail -3 gpurun_out/fix_cfg5_nx158.err | cut -c1-300
