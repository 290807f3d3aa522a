#!/bin/bash
# round-2 GPU call L: polyphase stride-2 weight gradient, warp-per-pixel ToRGB forward for wide layers
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/l_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
echo "bench rc=$?" >> gpurun_out/l_bench.err
timeout 300 python scripts/timeline_step.py > gpurun_out/l_timeline.txt 2>&1
timeout 600 python bench.py --workload gd_step --steps 5 > gpurun_out/l_bench_gd.json 2> gpurun_out/l_bench_gd.err
timeout 300 python scripts/timeline_gd_step.py > gpurun_out/l_timeline_gd.txt 2>&1
tail -4 gpurun_out/l_pytest.log; grep -n "^FAILED" gpurun_out/l_pytest.log; tail -c 300 gpurun_out/l_bench.json; echo; tail -c 300 gpurun_out/l_bench_gd.json; echo; head -30 gpurun_out/l_timeline_gd.txt | tail -27
