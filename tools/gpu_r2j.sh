#!/bin/bash
# 8-GPU pass: cfg3 at its named size (16384 scenarios = 2048 per GPU), then cfg2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --config cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_cfg3_8gpu.json 2> gpurun_out/r2j_cfg3_8gpu.err
echo "rc=$?" >> gpurun_out/r2j_cfg3_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_cfg2_8gpu.json 2> gpurun_out/r2j_cfg2_8gpu.err
echo "rc=$?" >> gpurun_out/r2j_cfg2_8gpu.err
tail -c 600 gpurun_out/r2j_cfg3_8gpu.json; tail -3 gpurun_out/r2j_cfg3_8gpu.err; tail -c 400 gpurun_out/r2j_cfg2_8gpu.json
