#!/bin/bash
mkdir -p gpurun_out
timeout 280 python -m pytest tests/test_multi_gpu.py tests/test_two_ranks_one_gpu.py -m gpu -x -q -s 2>&1 | tail -6 > gpurun_out/s17_multi_gpu_pytest.log; cat gpurun_out/s17_multi_gpu_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/s17_weak2_pack.json 2> gpurun_out/s17_weak2_pack.err
B200FLOW_PACK_U16=0 timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/s17_weak2_nopack.json 2> gpurun_out/s17_weak2_nopack.err
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --scaling strong > gpurun_out/s17_strong2_pack.json 2> gpurun_out/s17_strong2_pack.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s17_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'N', d['n_gpus'], d['scaling'], 'ms/step %.2f'%d['ms_per_step'], 'value %.1f M/s'%(d['value']/1e6), 'hash', d['forest_hash'], 'exch', d.get('level_exchange_ms'))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-800:])
PY
