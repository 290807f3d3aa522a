#!/bin/bash
# round-2 GPU call V (4 GPUs): the headline line launched exactly as the driver's scaling run issues it at N = 4
mkdir -p gpurun_out
SECONDS=0
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/v_bench4.json 2> gpurun_out/v_bench4.err
echo "bench4 rc=$? after ${SECONDS}s"
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/v_bench4.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['n_gpus'], d['config'].get('gradient_all_reduce'), d['e2e']['value'], (d.get('tf32x3') or {}).get('value'))
except Exception as e:
    print('no line', e)
PY
