#!/bin/bash
# round-1 closing GPU call: headline bench, full GPU suite, smoke, configs[2] bench, G+D step kernel breakdown
mkdir -p gpurun_out
timeout 120 python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
timeout 300 python -m pytest tests/ -m gpu -q > gpurun_out/f_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/f_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
timeout 100 python bench.py --workload gd_step --steps 5 --warmup 3 > gpurun_out/f_bench_gd.json 2> gpurun_out/f_bench_gd.err
timeout 100 python scripts/timeline_gd_step.py > gpurun_out/f_gd_timeline.txt 2>&1
grep -E "^FAILED|passed|failed|exit" gpurun_out/f_tests.log | tail -8; tail -1 gpurun_out/f_smoke.log; cut -c1-200 gpurun_out/f_bench.json; cut -c1-200 gpurun_out/f_bench_gd.json; head -12 gpurun_out/f_gd_timeline.txt
