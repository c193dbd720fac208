#!/bin/bash
# ordering experiment: SuperLU-MMD (default) against the multi-start AMD candidates, cfg2 and cfg3
mkdir -p gpurun_out
O='--no-cpu-baseline --no-e2e --steps 3 --warmup 3'
python bench.py $O > gpurun_out/ord_cfg2_default.json 2> gpurun_out/ord_cfg2_default.err
python bench.py $O --lu-opt '{"orderings":["nd","mmd","amd"]}' > gpurun_out/ord_cfg2_amd.json 2> gpurun_out/ord_cfg2_amd.err
python bench.py $O --config cfg3 --lu-opt '{"orderings":["nd","mmd","amd"]}' > gpurun_out/ord_cfg3_amd.json 2> gpurun_out/ord_cfg3_amd.err
tail -n 2 gpurun_out/ord_*.json gpurun_out/ord_*.err
