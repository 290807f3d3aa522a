#!/bin/bash
# round-2 GPU call H: suite with CTA-pair wgrad + split-K heuristic, micro-bench, step bench, gd_step bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest.log
timeout 300 python scripts/bench_conv.py > gpurun_out/h_bench_conv.jsonl 2> gpurun_out/h_bench_conv.err
SGV_WGRAD_PAIR=0 timeout 200 python scripts/bench_conv.py main4 > gpurun_out/h_bench_conv_nopair.jsonl 2>> gpurun_out/h_bench_conv.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
echo "bench rc=$?" >> gpurun_out/h_bench.err
timeout 600 python bench.py --workload gd_step --steps 5 > gpurun_out/h_bench_gd.json 2> gpurun_out/h_bench_gd.err
timeout 300 python scripts/timeline_gd_step.py > gpurun_out/h_timeline_gd.txt 2>&1
tail -4 gpurun_out/h_pytest.log
