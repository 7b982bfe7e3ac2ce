#!/bin/bash
# usage: bash tools/gpu_scale_kdd.sh N   — the two KDD99-full lines (weak, strong) on N GPUs of one box
N=$1; O=gpurun_out/scale; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > $O/kdd_full_weak_n$N.json 2> $O/kdd_full_weak_n$N.err
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --scaling strong > $O/kdd_full_strong_n$N.json 2> $O/kdd_full_strong_n$N.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/scale/kdd_full_*_n$N.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'N', d['n_gpus'], d.get('scaling'), 'ms/step %.2f'%d['ms_per_step'], 'value %.1f M/s'%(d['value']/1e6), 'hash', d.get('forest_hash'), 'exch', d.get('level_exchange_ms'))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-500:])
PY
