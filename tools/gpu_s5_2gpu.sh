#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s5_gpus.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > gpurun_out/s5_multi_gpu_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s5_multi_gpu_pytest.log
tail -4 gpurun_out/s5_multi_gpu_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/s5_weak2.json 2> gpurun_out/s5_weak2.err
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --scaling strong > gpurun_out/s5_strong2.json 2> gpurun_out/s5_strong2.err
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --scaling strong > gpurun_out/s5_strong1.json 2> gpurun_out/s5_strong1.err
B200FLOW_RS_CHUNKS=1 timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/s5_weak2_chunks1.json 2> gpurun_out/s5_weak2_chunks1.err
timeout 400 $TR bench.py --gpus 2 --steps 5 --warmup 2 --no-e2e --workload cicids_full --scaling strong > gpurun_out/s5_cicids_strong2.json 2> gpurun_out/s5_cicids_strong2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s5_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'N', d['n_gpus'], d['scaling'], 'ms/step %.2f'%d['ms_per_step'], 'value %.1f M/s'%(d['value']/1e6), 'hash', d['forest_hash'], 'nodes', d['forest_nodes'], 'exch', d.get('level_exchange_ms'))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-800:])
PY
