#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_networks_gpu.py tests/test_train_step_gpu.py tests/test_dropin_gpu.py -q -s > gpurun_out/c_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c_tests.log
timeout 200 python bench.py --workload gd_step --steps 5 --warmup 3 --fused-d 1 > gpurun_out/c_bench_gd_fused.json 2> gpurun_out/c_bench_gd_fused.err
timeout 200 python bench.py --workload gd_step --steps 5 --warmup 3 --fused-d 0 > gpurun_out/c_bench_gd_unfused.json 2> gpurun_out/c_bench_gd_unfused.err
grep -E "^FAILED|passed|failed" gpurun_out/c_tests.log; cut -c1-260 gpurun_out/c_bench_gd_fused.json; cut -c1-260 gpurun_out/c_bench_gd_unfused.json; tail -2 gpurun_out/c_bench_gd_fused.err
