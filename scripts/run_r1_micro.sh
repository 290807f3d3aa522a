#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_networks_gpu.py tests/test_train_step_gpu.py -q > gpurun_out/m_tests.log 2>&1
echo "exit $?" >> gpurun_out/m_tests.log
timeout 40 python bench.py --workload gd_step --steps 5 --warmup 3 > gpurun_out/m_bench_gd.json 2> gpurun_out/m_bench_gd.err
grep -E "^FAILED|passed|failed|exit" gpurun_out/m_tests.log; cut -c1-200 gpurun_out/m_bench_gd.json
