#!/bin/bash
# round-2 GPU session 2: encode_bins v2 check, launch-shape sweep, ncu captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_records.py -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log
tail -3 gpurun_out/s2_pytest.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
for sh in default 8x1 16x2 16x1 32x1; do
  if [ $sh = default ]; then unset B200FLOW_ROUTE_SHAPE; else export B200FLOW_ROUTE_SHAPE=$sh; fi
  timeout 200 python bench.py --workload kdd_full $B > gpurun_out/s2_kdd_full_$sh.json 2> gpurun_out/s2_kdd_full_$sh.err
done
for sh in default 8x2 16x1 16x2; do
  if [ $sh = default ]; then unset B200FLOW_ROUTE_SHAPE; else export B200FLOW_ROUTE_SHAPE=$sh; fi
  timeout 200 python bench.py --workload cicids_wed $B > gpurun_out/s2_cicids_wed_$sh.json 2> gpurun_out/s2_cicids_wed_$sh.err
done
for sh in default 8x2 8x1 32x1; do
  if [ $sh = default ]; then unset B200FLOW_ROUTE_SHAPE; else export B200FLOW_ROUTE_SHAPE=$sh; fi
  timeout 200 python bench.py --workload kdd_full --classes 23 $B > gpurun_out/s2_kdd23_$sh.json 2> gpurun_out/s2_kdd23_$sh.err
done
unset B200FLOW_ROUTE_SHAPE
timeout 300 python bench.py --workload cicids_full $B > gpurun_out/s2_cicids_full.json 2> gpurun_out/s2_cicids_full.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'chunk', d.get('route_chunk'), {kk:(round(v['ms_per_step'],2), v['launches_per_step']) for kk,v in k.items() if kk in ('route_hist_level','encode_bins','score_level','hist_level','predict')})
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-400:])
PY
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:route_hist_level -s 9 -c 1 -o gpurun_out/s2_route_kdd python bench.py --workload kdd_full --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/s2_ncu_route_kdd.log 2>&1
timeout 400 $NCU -k regex:encode_bins -s 0 -c 2 -o gpurun_out/s2_bins_kdd python bench.py --workload kdd_full --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/s2_ncu_bins_kdd.log 2>&1
timeout 400 $NCU -k regex:"route_hist_level|score_level" -s 18 -c 2 -o gpurun_out/s2_route_score_cicids python bench.py --workload cicids_wed --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/s2_ncu_cicids.log 2>&1
timeout 300 $NCU -k regex:encode_bins -s 0 -c 1 -o gpurun_out/s2_bins_cicids python bench.py --workload cicids_wed --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/s2_ncu_bins_cicids.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
