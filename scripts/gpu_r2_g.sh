#!/bin/bash
# round-2 GPU call G: split-K cluster reduction of the per-tap conv kernel (small planes): suite + micro-bench + step bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/g_pytest.log
for name in b4.conv1 b8.conv1 b16.conv1; do
  SGV_CONV_SPLITK=1 timeout 120 python scripts/bench_conv.py $name >> gpurun_out/g_small_split.jsonl 2>> gpurun_out/g_small.err
  SGV_CONV_SPLITK=0 timeout 120 python scripts/bench_conv.py $name >> gpurun_out/g_small_nosplit.jsonl 2>> gpurun_out/g_small.err
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-second-mode > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "bench rc=$?" >> gpurun_out/g_bench.err
tail -4 gpurun_out/g_pytest.log
