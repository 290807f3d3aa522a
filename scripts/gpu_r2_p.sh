#!/bin/bash
# round-2 GPU call P (8 GPUs): headline workload (early gradient bucket inside the graph) and the G + D step (one graph per phase) over NCCL
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:2}"; }
SECONDS=0
timeout 170 bash -c "$(declare -f run); run 29551 --steps 20 --warmup 5 --no-second-mode" > gpurun_out/p_bench8.json 2> gpurun_out/p_bench8.err; echo "bench8 rc=$? after ${SECONDS}s"
timeout 150 bash -c "$(declare -f run); run 29552 --workload gd_step --steps 10 --warmup 3" > gpurun_out/p_bench8_gd.json 2> gpurun_out/p_bench8_gd.err; echo "gd8 rc=$? after ${SECONDS}s"
for f in p_bench8 p_bench8_gd; do python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', d['value'], d['ms_per_step'], d['n_gpus'], d['config'].get('gradient_all_reduce'), d['config'].get('cuda_graph'), d['e2e']['value'], d.get('clocks'))
except Exception as e:
    print('$f', 'no line', e)
PY
done
grep -v "^\*\|OMP_NUM\|^W0\|^$\|UserWarning\|run_backward" gpurun_out/p_bench8.err | tail -6
