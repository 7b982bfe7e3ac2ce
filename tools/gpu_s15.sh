#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_fused_records.py -m gpu -x -q > gpurun_out/s15_pytest.log 2>&1; tail -2 gpurun_out/s15_pytest.log
B="--steps 8 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 200 python bench.py --workload kdd_full $B > gpurun_out/s15_kdd_full.json 2> gpurun_out/s15_kdd_full.err
timeout 200 python bench.py --workload cicids_full --trees 100 --depth 16 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/s15_cicids_deep.json 2> gpurun_out/s15_cicids_deep.err
timeout 200 python bench.py --workload kdd_script $B > gpurun_out/s15_kdd_script.json 2> gpurun_out/s15_kdd_script.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s15_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'route', round(d['kernels']['route_hist_level']['ms_per_step'],2))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:route_hist_level -s 9 -c 2 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e 2>&1 | grep -E "route_hist|dram__|lts__|gpu__time" | head -12
