#!/bin/bash
O=gpurun_out/final; mkdir -p $O
for wl in kdd_script cicids_full cicids_wed cicids_script kdd10; do
  timeout 120 python bench.py --workload $wl --steps 10 --warmup 3 > $O/bench_$wl.json 2> $O/bench_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open('$O/bench_$wl.json').read().strip().splitlines()[-1]); c=d.get('cpu_baseline') or {}
    print('$wl', 'ms/step %.2f'%d['ms_per_step'], 'value %.1f M/s'%(d['value']/1e6), 'e2e %.2f'%d['e2e']['ms_per_step'], 'route %.3f'%d['kernels']['route_hist_level']['ms_per_step'], c.get('labels_equal'), c.get('forest_equal'))
except Exception as e: print('$wl ERR', e)
PY
done
