#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/debug_d_layers.py > gpurun_out/b_debug_d.log 2>&1
timeout 300 python -m pytest tests/test_train_aux_gpu.py tests/test_networks_gpu.py tests/test_train_step_gpu.py -q -s > gpurun_out/b_new_tests.log 2>&1
echo "new tests exit $?" >> gpurun_out/b_new_tests.log
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
echo "bench exit $?" >> gpurun_out/b_bench.err
grep -E "rel_err|logits" gpurun_out/b_debug_d.log | head -70; grep -E "^FAILED|passed|failed|fused adam" gpurun_out/b_new_tests.log; cut -c1-300 gpurun_out/b_bench.json
