#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_fused_records.py tests/test_golden.py tests/test_pyspark_shim.py -m gpu -x -q > gpurun_out/s12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s12_pytest.log
tail -3 gpurun_out/s12_pytest.log
S="--workload stream --rows 16777216 --steps 4 --warmup 2 --no-e2e"
B200FLOW_PRED_MODE=0 timeout 300 python bench.py $S > gpurun_out/s12_stream_mode0.json 2> gpurun_out/s12_stream_mode0.err
for kb in 26 52 104; do
  B200FLOW_PRED_TOP_KB=$kb timeout 300 python bench.py $S > gpurun_out/s12_stream_walk_$kb.json 2> gpurun_out/s12_stream_walk_$kb.err
done
B200FLOW_PRED_TOP_KB=52 B200FLOW_PRED_ROWS=4 timeout 300 python bench.py $S > gpurun_out/s12_stream_walk_52_r4.json 2> gpurun_out/s12_stream_walk_52_r4.err
K="--workload kdd_full --steps 6 --warmup 3 --no-e2e --no-cpu-baseline"
B200FLOW_PRED_MODE=0 timeout 300 python bench.py $K > gpurun_out/s12_kdd_mode0.json 2> gpurun_out/s12_kdd_mode0.err
timeout 300 python bench.py $K > gpurun_out/s12_kdd_walk.json 2> gpurun_out/s12_kdd_walk.err
timeout 300 python bench.py --workload kdd_script --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/s12_kdd_script_walk.json 2> gpurun_out/s12_kdd_script_walk.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s12_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'value %.1f M/s'%(d['value']/1e6), 'predict', round(k['predict']['ms_per_step'],2))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
