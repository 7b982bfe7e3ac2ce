#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_records.py tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s7_pytest.log
tail -3 gpurun_out/s7_pytest.log
B="--steps 8 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 200 python bench.py --workload kdd_full $B > gpurun_out/s7_kdd_full.json 2> gpurun_out/s7_kdd_full.err
B200FLOW_ROUTE_LB3=1 timeout 200 python bench.py --workload kdd_full $B > gpurun_out/s7_kdd_full_lb3.json 2> gpurun_out/s7_kdd_full_lb3.err
for wl in cicids_wed cicids_full; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-sklearn > gpurun_out/s7_$wl.json 2> gpurun_out/s7_$wl.err
done
timeout 300 python tools/bench_encode.py --iters 20 > gpurun_out/s7_bench_encode.txt 2>&1
timeout 300 python bench.py --workload stream --rows 16777216 --steps 4 --warmup 2 > gpurun_out/s7_stream_2e24.json 2> gpurun_out/s7_stream_2e24.err
timeout 400 python bench.py --workload stream --steps 5 --warmup 2 > gpurun_out/s7_stream.json 2> gpurun_out/s7_stream.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s7_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']; c=d.get('cpu_baseline') or {}
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'value %.1f M/s'%(d['value']/1e6), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],2), 'roof', round(d['roofline']['frac'] or 0,3), {kk:(round(v['ms_per_step'],2), v['launches_per_step']) for kk,v in k.items() if v['ms_per_step']>0.3}, {k2:c.get(k2) for k2 in ('value','labels_equal','forest_equal')})
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
cat gpurun_out/s7_bench_encode.txt | cut -c1-250
