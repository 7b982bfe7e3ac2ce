#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_encode.py --iters 20 > gpurun_out/s8_bench_encode.txt 2>&1
B200FLOW_ENC_SMEM_KB=100 timeout 300 python tools/bench_encode.py --iters 20 > gpurun_out/s8_bench_encode_100kb.txt 2>&1
S="--workload stream --rows 16777216 --steps 4 --warmup 2 --no-e2e"
timeout 300 python bench.py $S > gpurun_out/s8_stream_r2.json 2> gpurun_out/s8_stream_r2.err
B200FLOW_PRED_ROWS=4 timeout 300 python bench.py $S > gpurun_out/s8_stream_r4.json 2> gpurun_out/s8_stream_r4.err
B200FLOW_TOP_LEVELS=10 timeout 300 python bench.py $S > gpurun_out/s8_stream_top10.json 2> gpurun_out/s8_stream_top10.err
B200FLOW_TOP_LEVELS=10 B200FLOW_PRED_ROWS=4 timeout 300 python bench.py $S > gpurun_out/s8_stream_r4_top10.json 2> gpurun_out/s8_stream_r4_top10.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s8_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'value %.1f M/s'%(d['value']/1e6), {kk:(round(v['ms_per_step'],2), v['launches_per_step']) for kk,v in k.items()})
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
for f in gpurun_out/s8_bench_encode.txt gpurun_out/s8_bench_encode_100kb.txt; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('  ',d['plan'], 'ms %.3f'%d['ms'], 'frac %.3f'%d['frac'])
PY
done
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:predict_kernel -c 1 -o gpurun_out/s8_predict_stream python bench.py $S --steps 1 --warmup 0 > gpurun_out/s8_ncu_predict.log 2>&1
timeout 300 $NCU -k regex:encode_kernel -c 1 -o gpurun_out/s8_encode python tools/bench_encode.py --iters 1 > gpurun_out/s8_ncu_encode.log 2>&1
