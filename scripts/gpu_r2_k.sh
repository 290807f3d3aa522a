#!/bin/bash
# round-2 GPU call K: demod kernels, up-2 FIR quad kernel, split-K heuristic (KS <= 4), EqualizedLinear routing fix
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/k_pytest.log
timeout 120 python scripts/bench_conv.py small > gpurun_out/k_bench_conv_small.jsonl 2> gpurun_out/k_bench_conv_small.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
echo "bench rc=$?" >> gpurun_out/k_bench.err
timeout 300 python scripts/timeline_step.py > gpurun_out/k_timeline.txt 2>&1
timeout 600 python bench.py --workload gd_step --steps 5 > gpurun_out/k_bench_gd.json 2> gpurun_out/k_bench_gd.err
timeout 300 python scripts/timeline_gd_step.py > gpurun_out/k_timeline_gd.txt 2>&1
timeout 300 python scripts/bench_ops.py > gpurun_out/k_bench_ops.jsonl 2> gpurun_out/k_bench_ops.err
tail -4 gpurun_out/k_pytest.log; grep -n "^FAILED" gpurun_out/k_pytest.log; tail -c 400 gpurun_out/k_bench.json; echo; tail -c 300 gpurun_out/k_bench_gd.json; echo; head -12 gpurun_out/k_timeline_gd.txt | tail -9
