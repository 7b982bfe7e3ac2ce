#!/bin/bash
set -u
O=gpurun_out/final; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:route_hist_level -s 3 -c 1 -o $O/ncu_route_cicids_full python bench.py --workload cicids_full --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/ncu_route_cicids_full.log 2>&1
timeout 400 $NCU -k regex:route_hist_level -s 3 -c 1 -o $O/ncu_route_kdd_script python bench.py --workload kdd_script --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/ncu_route_kdd_script.log 2>&1
timeout 400 $NCU -k regex:"csv_rows_kernel|csv_count|csv_line" -c 8 -o $O/ncu_csv python tools/bench_csv.py 300000 > $O/ncu_csv.log 2>&1
# shape re-check for the rotated update
for s in 8x2 8x1 16x2; do B200FLOW_ROUTE_SHAPE=$s timeout 300 python bench.py --workload kdd_full --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $O/shape_kdd_full_$s.json 2>/dev/null; done
for s in 8x1 8x2 16x1; do B200FLOW_ROUTE_SHAPE=$s timeout 300 python bench.py --workload cicids_wed --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $O/shape_cicids_wed_$s.json 2>/dev/null; done
for s in 16x1 32x1 8x1; do B200FLOW_ROUTE_SHAPE=$s timeout 300 python bench.py --workload kdd_script --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $O/shape_kdd_script_$s.json 2>/dev/null; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/final/shape_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f.split('/')[-1], round(d['ms_per_step'],3), round(d['kernels']['route_hist_level']['ms_per_step'],3))
    except Exception as e: print(f, 'ERR', e)
PY
