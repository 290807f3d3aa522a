#!/bin/bash
# round-2 GPU call O (2 GPUs): gradient all-reduce of the conv-weight bucket captured inside the step's graph (under the backward tail) vs one
# collective after the replay
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_synthesis_gpu.py tests/test_dense_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/o_pytest.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
timeout 500 bash -c "$(declare -f run); run 29531 --steps 20 --warmup 5 --no-second-mode" > gpurun_out/o_bench2_overlap.json 2> gpurun_out/o_bench2_overlap.err; echo "overlap rc=$?"
timeout 500 bash -c "$(declare -f run); run 29532 --steps 20 --warmup 5 --no-second-mode --no-overlap" > gpurun_out/o_bench2_plain.json 2> gpurun_out/o_bench2_plain.err; echo "plain rc=$?"
for f in o_bench2_overlap o_bench2_plain; do python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', d['value'], d['ms_per_step'], d['config'].get('gradient_all_reduce'), d['e2e']['value'])
except Exception as e:
    print('$f', 'no line', e)
PY
done
grep -v "^\*\|OMP_NUM\|^W0\|^$" gpurun_out/o_bench2_overlap.err | tail -12
