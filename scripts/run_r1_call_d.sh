#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/debug_d_grads.py > gpurun_out/d_debug_grads.log 2>&1
timeout 300 python -m pytest tests/test_networks_gpu.py tests/test_train_step_gpu.py -q -s > gpurun_out/d_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/d_tests.log
grep -E "^FAILED|passed|failed" gpurun_out/d_tests.log; grep -c . gpurun_out/d_debug_grads.log
