#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s18_pytest.log; tail -4 gpurun_out/s18_pytest.log
B="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
for wl in kdd_full kdd_script cicids_full kdd10; do
timeout 300 python bench.py --workload $wl $B > gpurun_out/s18_$wl.json 2> gpurun_out/s18_$wl.err
done
timeout 200 python tools/timeline.py --workload kdd_full 2>/dev/null | grep -E "^span|ms by gap" 
timeout 200 python tools/timeline.py --workload kdd_script 2>/dev/null | grep -E "^span|ms by gap" 
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s18_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'launches', d['gpu_launches'])
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
