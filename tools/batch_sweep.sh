#!/bin/bash
# per-launch kernel times vs batch size (footprint effects: TLB reach, L2 residency)
for b in 148 592 1184 2368 4096 8192; do
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --batch $b 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['classes']; B=d['config']['batch_per_gpu']
print('B %5d it/s %8.0f | us per scenario: factor %.2f solve %.2f J %.3f F %.3f'%(B,d['value'],1e3*r['lu_refactor']['ms_per_launch']/B,1e3*r['lu_solve']['ms_per_launch']/B,1e3*r['J']['ms_per_launch']/B,1e3*r['F']['ms_per_launch']/B))"
done
