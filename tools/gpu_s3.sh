#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_records.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.log
tail -3 gpurun_out/s3_pytest.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
for wl in kdd_full cicids_wed cicids_full; do
  timeout 200 python bench.py --workload $wl $B > gpurun_out/s3_$wl.json 2> gpurun_out/s3_$wl.err
done
timeout 200 python tools/timeline.py --workload kdd_full > gpurun_out/s3_timeline_kdd.txt 2>&1
timeout 200 python tools/timeline.py --workload kdd_script > gpurun_out/s3_timeline_kdd_script.txt 2>&1
timeout 200 python tools/profile_levels.py --workload kdd_full > gpurun_out/s3_levels_kdd.txt 2>&1
timeout 200 python tools/profile_levels.py --workload cicids_full > gpurun_out/s3_levels_cicids_full.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s3_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']
        print(f, 'ms/step %.2f'%d['ms_per_step'], {kk:(round(v['ms_per_step'],2), v['launches_per_step']) for kk,v in k.items()})
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-400:])
PY
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:encode_bins -s 0 -c 1 -o gpurun_out/s3_bins_kdd python bench.py --workload kdd_full --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/s3_ncu_bins_kdd.log 2>&1
timeout 400 $NCU -k regex:"bag_weights|predict_kernel|dedup|group_" -s 0 -c 9 -o gpurun_out/s3_misc_kdd python bench.py --workload kdd_full --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/s3_ncu_misc_kdd.log 2>&1
