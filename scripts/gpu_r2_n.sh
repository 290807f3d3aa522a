#!/bin/bash
# round-2 GPU call N (2 GPUs): synthesis and G+D step over NCCL; the G+D step replays one CUDA graph per phase with the all-reduce + update between
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
timeout 600 bash -c "$(declare -f run); run 29521 --steps 20 --warmup 5 --no-second-mode" > gpurun_out/n_bench2.json 2> gpurun_out/n_bench2.err; echo "bench2 rc=$?"
timeout 600 bash -c "$(declare -f run); run 29522 --workload gd_step --steps 10 --warmup 3" > gpurun_out/n_bench2_gd.json 2> gpurun_out/n_bench2_gd.err; echo "gd2 rc=$?"
tail -c 500 gpurun_out/n_bench2.json; echo; tail -c 700 gpurun_out/n_bench2_gd.json; echo; grep -v "^\*\|OMP_NUM" gpurun_out/n_bench2_gd.err | tail -8
