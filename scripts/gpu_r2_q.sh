#!/bin/bash
# round-2 GPU call Q (2 GPUs): last check of the tree as it stands — smoke + full suite on GPU 0, headline line at N = 1, then the N = 2 launch
# exactly as the driver issues it (early gradient bucket through FlatModuleState.begin_backward / finish_backward)
mkdir -p gpurun_out
SECONDS=0
timeout 200 python __graft_entry__.py smoke > gpurun_out/q_smoke.log 2>&1; echo "smoke rc=$? after ${SECONDS}s"; tail -1 gpurun_out/q_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$? after ${SECONDS}s"; tail -2 gpurun_out/q_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/q_bench1.json 2> gpurun_out/q_bench1.err; echo "bench1 rc=$? after ${SECONDS}s"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/q_bench2.json 2> gpurun_out/q_bench2.err; echo "bench2 rc=$? after ${SECONDS}s"
for f in q_bench1 q_bench2; do python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', d['value'], d['ms_per_step'], d['n_gpus'], d['config'].get('gradient_all_reduce'), d['e2e']['value'], (d.get('tf32x3') or {}).get('value'))
except Exception as e:
    print('$f', 'no line', e)
PY
done
