#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_fused_records.py tests/test_golden.py -m gpu -x -q > gpurun_out/s9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s9_pytest.log
tail -3 gpurun_out/s9_pytest.log
timeout 200 python bench.py --workload kdd_full --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s9_kdd_full.json 2> gpurun_out/s9_kdd_full.err
timeout 200 python bench.py --workload kdd_script --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s9_kdd_script.json 2> gpurun_out/s9_kdd_script.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s9_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['kernels']
        print(f, 'ms/step %.2f'%d['ms_per_step'], {kk:(round(v['ms_per_step'],2), v['launches_per_step']) for kk,v in k.items()})
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
for kb in 40 52 64 76 100 150; do
  B200FLOW_ENC_SMEM_KB=$kb timeout 200 python tools/bench_encode.py --iters 20 > gpurun_out/s9_enc_$kb.txt 2>&1
  echo "ENC_SMEM_KB=$kb"; python - gpurun_out/s9_enc_$kb.txt <<'PY'
import json,sys
print('   ', ' | '.join('%s %.3f'%(json.loads(l)['plan'][:18], json.loads(l)['frac']) for l in open(sys.argv[1]) if l.startswith('{')))
PY
done
