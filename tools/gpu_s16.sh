#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s16_pytest.log; tail -4 gpurun_out/s16_pytest.log
B="--steps 8 --warmup 3 --no-e2e --no-sklearn"
timeout 300 python bench.py --workload kdd_full $B > gpurun_out/s16_kdd_full.json 2> gpurun_out/s16_kdd_full.err
B200FLOW_PACK_U16=0 timeout 300 python bench.py --workload kdd_full $B --no-cpu-baseline > gpurun_out/s16_kdd_full_nopack.json 2> gpurun_out/s16_kdd_full_nopack.err
timeout 400 python bench.py --workload cicids_full --trees 100 --depth 16 --steps 3 --warmup 2 --no-e2e --no-sklearn > gpurun_out/s16_cicids_deep.json 2> gpurun_out/s16_cicids_deep.err
timeout 300 python bench.py --workload kdd_script $B > gpurun_out/s16_kdd_script.json 2> gpurun_out/s16_kdd_script.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s16_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d.get('cpu_baseline') or {}
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'route', round(d['kernels']['route_hist_level']['ms_per_step'],2), 'score', round(d['kernels']['score_level']['ms_per_step'],2), 'parity', c.get('labels_equal'), c.get('forest_equal'))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
