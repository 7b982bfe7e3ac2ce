#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4_pytest.log
tail -3 gpurun_out/s4_pytest.log
timeout 600 python bench.py --workload kdd_full --steps 10 --warmup 3 > gpurun_out/s4_kdd_full.json 2> gpurun_out/s4_kdd_full.err
timeout 300 python bench.py --workload kdd_script --steps 5 --warmup 3 --no-sklearn > gpurun_out/s4_kdd_script.json 2> gpurun_out/s4_kdd_script.err
timeout 300 python bench.py --workload cicids_script --steps 5 --warmup 3 --no-sklearn > gpurun_out/s4_cicids_script.json 2> gpurun_out/s4_cicids_script.err
timeout 300 python bench.py --workload kdd10 --steps 5 --warmup 3 --no-sklearn > gpurun_out/s4_kdd10.json 2> gpurun_out/s4_kdd10.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s4_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d.get('cpu_baseline') or {}
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],2), 'launches', d['gpu_launches'], 'cpu', {k:c.get(k) for k in ('value','labels_equal','forest_equal','macro_f1_equals_gpu')}, (c.get('sklearn') or {}).get('value'))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
