#!/bin/bash
# round-2 GPU session 1: tests + first bench lines of every workload
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
tail -5 gpurun_out/s1_pytest.log
for wl in kdd_full; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s1_${wl}_records.json 2> gpurun_out/s1_${wl}_records.err
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --path dense > gpurun_out/s1_${wl}_dense.json 2> gpurun_out/s1_${wl}_dense.err
done
for wl in kdd_script cicids_wed cicids_full cicids_script kdd10; do
  timeout 300 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline --no-sklearn > gpurun_out/s1_${wl}.json 2> gpurun_out/s1_${wl}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s1_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'e2e', d['e2e'] and '%.2f'%d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'chunk', d.get('route_chunk'), 'passes', d.get('route_passes'))
        print('   ', {kk:(round(v['ms_per_step'],3), v['launches_per_step']) for kk,v in k.items()})
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
