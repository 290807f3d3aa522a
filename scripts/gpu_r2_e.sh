#!/bin/bash
# round-2 GPU call E: GPU suite (CTA pairs on by default, stacked-M wgrad), fused-vs-unfused D localisation, conv / wgrad micro-bench, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e_pytest.log
timeout 300 python scripts/debug_d128.py > gpurun_out/e_debug_d128.txt 2>&1
timeout 300 python scripts/bench_conv.py main4 > gpurun_out/e_bench_conv.jsonl 2> gpurun_out/e_bench_conv.err
SGV_WGRAD_S64=0 timeout 300 python scripts/bench_conv.py b256.conv1 > gpurun_out/e_bench_conv_nos64.jsonl 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
echo "bench rc=$?" >> gpurun_out/e_bench.err
timeout 600 python bench.py --workload gd_step --steps 5 > gpurun_out/e_bench_gd.json 2> gpurun_out/e_bench_gd.err
tail -4 gpurun_out/e_pytest.log
