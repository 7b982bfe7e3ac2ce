#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s10_pytest.log
tail -15 gpurun_out/s10_pytest.log
timeout 300 python bench.py --workload kdd_full --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/s10_kdd_full.json 2> gpurun_out/s10_kdd_full.err
timeout 300 python bench.py --workload kdd_script --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/s10_kdd_script.json 2> gpurun_out/s10_kdd_script.err
timeout 300 python bench.py --workload kdd10 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/s10_kdd10.json 2> gpurun_out/s10_kdd10.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s10_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'e2e', d['e2e'] and (round(d['e2e']['ms_per_step'],2), d['e2e']['macro_f1_equals_resident']), {kk:(round(v['ms_per_step'],2), v['launches_per_step']) for kk,v in k.items() if v['ms_per_step']>0.25})
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
