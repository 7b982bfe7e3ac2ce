#!/bin/bash
O=gpurun_out/scale; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544"
timeout 300 $TR bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > $O/kdd_full_weak_n4_b.json 2> $O/kdd_full_weak_n4_b.err
B200FLOW_RS_MIN_BYTES=1000000000000 timeout 300 $TR bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > $O/kdd_full_weak_n4_allreduce.json 2> $O/kdd_full_weak_n4_allreduce.err
NCCL_DEBUG=INFO timeout 300 $TR bench.py --gpus 4 --steps 2 --warmup 1 --no-e2e > $O/kdd_full_weak_n4_dbg.json 2> $O/kdd_full_weak_n4_dbg.err
grep -i "NVLS\|algo\|Connected all\|channels" $O/kdd_full_weak_n4_dbg.err | head -12
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/scale/kdd_full_weak_n4_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms/step %.2f'%d['ms_per_step'], 'exch', d.get('level_exchange_ms'), 'score', round(d['kernels']['score_level']['ms_per_step'],2))
    except Exception as e:
        print(f,'ERR',e)
PY
