#!/bin/bash
# round-2 GPU call C: full GPU suite, smoke, bench, gd_step bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/c_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/c_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
echo "bench rc=$?" >> gpurun_out/c_bench.err
timeout 600 python bench.py --workload gd_step --steps 5 > gpurun_out/c_bench_gd.json 2> gpurun_out/c_bench_gd.err
echo "gd rc=$?" >> gpurun_out/c_bench_gd.err
timeout 300 python scripts/timeline_gd_step.py > gpurun_out/c_timeline_gd.txt 2>&1
tail -4 gpurun_out/c_pytest.log
