#!/bin/bash
# round-2 GPU call J: reduce-scatter split-K + exact-fp32 dense kernels: suite, small-plane sweep, dense micro-bench, step bench + timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j_pytest.log
timeout 200 python scripts/bench_dense.py > gpurun_out/j_bench_dense.jsonl 2> gpurun_out/j_bench_dense.err
for cfg in "0 0" "256 8" "256 4" "128 8" "128 4" "64 8" "64 4" "128 2"; do
  set -- $cfg
  echo "{\"SGV_CONV_V1_BN\": $1, \"SGV_CONV_V1_KS\": $2}" >> gpurun_out/j_bench_conv_small.jsonl
  SGV_CONV_V1_BN=$1 SGV_CONV_V1_KS=$2 timeout 120 python scripts/bench_conv.py small >> gpurun_out/j_bench_conv_small.jsonl 2>> gpurun_out/j_bench_conv_small.err
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
echo "bench rc=$?" >> gpurun_out/j_bench.err
timeout 300 python scripts/timeline_step.py > gpurun_out/j_timeline.txt 2>&1
timeout 600 python bench.py --workload gd_step --steps 5 > gpurun_out/j_bench_gd.json 2> gpurun_out/j_bench_gd.err
tail -4 gpurun_out/j_pytest.log; cat gpurun_out/j_bench_dense.jsonl; tail -c 300 gpurun_out/j_bench_dense.err
