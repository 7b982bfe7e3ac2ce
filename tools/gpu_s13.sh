#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_fused_records.py tests/test_golden.py tests/test_pyspark_shim.py -m gpu -x -q > gpurun_out/s13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s13_pytest.log
tail -12 gpurun_out/s13_pytest.log
S="--workload stream --rows 16777216 --steps 4 --warmup 2 --no-e2e"
B200FLOW_PRED_TREE_MAJOR=0 timeout 300 python bench.py $S > gpurun_out/s13_stream_old.json 2> gpurun_out/s13_stream_old.err
for kb in 40 80 120; do
  B200FLOW_PRED_TREE_KB=$kb timeout 300 python bench.py $S > gpurun_out/s13_stream_tm_$kb.json 2> gpurun_out/s13_stream_tm_$kb.err
done
K="--workload kdd_full --steps 6 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 python bench.py $K > gpurun_out/s13_kdd_tm.json 2> gpurun_out/s13_kdd_tm.err
timeout 300 python bench.py --workload kdd_script --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/s13_kdd_script_tm.json 2> gpurun_out/s13_kdd_script_tm.err
timeout 300 python bench.py --workload cicids_full --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/s13_cicids_full_tm.json 2> gpurun_out/s13_cicids_full_tm.err
timeout 300 python bench.py --workload cicids_full --trees 100 --depth 16 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/s13_cicids_deep_tm.json 2> gpurun_out/s13_cicids_deep_tm.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s13_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'value %.1f M/s'%(d['value']/1e6), 'predict', round(k['predict']['ms_per_step'],2))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
