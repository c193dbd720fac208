for opt in '{"n_warps":16,"n_groups":2}' '{"n_warps":8,"n_groups":1}' '{"n_warps":16,"n_groups":4}' '{"n_warps":32,"n_groups":4,"warp_front_max":16}' '{"n_warps":16,"n_groups":2,"warp_front_max":32}' '{"n_warps":8,"n_groups":1,"warp_front_max":32,"group_front_max":100}'; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --lu-opt "$opt" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['roofline']['classes']; print('$opt', round(d['value']), 'lu', round(c['lu_refactor']['ms_per_launch'],2), 'solve', round(c['lu_solve']['ms_per_launch'],2))"
done
