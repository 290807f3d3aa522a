#!/bin/bash
# call I: the N = 2 path on real NCCL (bench lines for the three workloads + the gloo/NCCL-agnostic tests)
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
timeout 600 bash -c "$(declare -f run); run 29511 --steps 10 --warmup 3 --no-second-mode" > gpurun_out/i_bench2.json 2> gpurun_out/i_bench2.err; echo "bench2 rc=$?"
timeout 600 bash -c "$(declare -f run); run 29512 --workload gd_step --steps 5 --warmup 3 --no-second-mode" > gpurun_out/i_bench2_gd.json 2> gpurun_out/i_bench2_gd.err; echo "gd2 rc=$?"
timeout 900 bash -c "$(declare -f run); run 29513 --workload full_loop --steps 4 --warmup 3 --no-second-mode" > gpurun_out/i_bench2_loop.json 2> gpurun_out/i_bench2_loop.err; echo "loop2 rc=$?"
timeout 600 bash -c "$(declare -f run); run 29514 --impl reference --steps 1 --warmup 0" > gpurun_out/i_ref2.json 2> gpurun_out/i_ref2.err; echo "ref2 rc=$?"
tail -c 600 gpurun_out/i_bench2.json; tail -c 400 gpurun_out/i_bench2_gd.json; tail -c 400 gpurun_out/i_bench2_loop.json; tail -c 300 gpurun_out/i_ref2.json
tail -5 gpurun_out/i_bench2.err gpurun_out/i_bench2_gd.err gpurun_out/i_bench2_loop.err
