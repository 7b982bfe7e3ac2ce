#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "encode" > gpurun_out/s14_pytest.log 2>&1; tail -2 gpurun_out/s14_pytest.log
for st in 3 2; do for kb in 44 52 60; do
  B200FLOW_ENC_STAGES=$st B200FLOW_ENC_SMEM_KB=$kb timeout 200 python tools/bench_encode.py --iters 20 > gpurun_out/s14_enc_${st}_$kb.txt 2>&1
  echo "stages=$st KB=$kb"; python - gpurun_out/s14_enc_${st}_$kb.txt <<'PY'
import json,sys
print('   ', ' | '.join('%s %.3f'%(json.loads(l)['plan'][:18], json.loads(l)['frac']) for l in open(sys.argv[1]) if l.startswith('{')))
PY
done; done
