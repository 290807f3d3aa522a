#!/bin/bash
# round-2 GPU call O2 (2 GPUs): the early-bucket run must EXIT cleanly (call O: result printed, ranks hung in the process-group teardown)
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
SECONDS=0
timeout 240 bash -c "$(declare -f run); run 29541 --steps 20 --warmup 5" > gpurun_out/o2_bench2_overlap.json 2> gpurun_out/o2_bench2_overlap.err; echo "overlap rc=$? after ${SECONDS}s"
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/o2_bench2_overlap.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['config'].get('gradient_all_reduce'), d['e2e']['value'], (d.get('tf32x3') or {}).get('value'))
except Exception as e:
    print('no line', e)
PY
grep -v "^\*\|OMP_NUM\|^W0\|^$" gpurun_out/o2_bench2_overlap.err | tail -6
