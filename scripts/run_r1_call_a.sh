#!/bin/bash
# one GPU call: new-kernel tests, the headline bench with the fused optimiser step, the configs[2] bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 300 python -m pytest tests/test_train_aux_gpu.py tests/test_networks_gpu.py tests/test_train_step_gpu.py -q -s > gpurun_out/a_new_tests.log 2>&1
echo "new tests exit $?" >> gpurun_out/a_new_tests.log
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench exit $?" >> gpurun_out/a_bench.err
timeout 240 python bench.py --steps 10 --warmup 3 --no-optimizer --no-cpu-baseline > gpurun_out/a_bench_noopt.json 2> gpurun_out/a_bench_noopt.err
timeout 240 python bench.py --workload gd_step --steps 5 --warmup 3 > gpurun_out/a_bench_gd.json 2> gpurun_out/a_bench_gd.err
echo "gd exit $?" >> gpurun_out/a_bench_gd.err
tail -3 gpurun_out/a_new_tests.log; cat gpurun_out/a_bench.json | cut -c1-400; cat gpurun_out/a_bench_gd.json | cut -c1-400; tail -3 gpurun_out/a_bench_gd.err
