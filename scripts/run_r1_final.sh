#!/bin/bash
# round-1 closing GPU call: headline bench, full GPU suite, smoke, the other BASELINE configs, G+D step kernel breakdown, ncu launch list
mkdir -p gpurun_out
timeout 120 python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
timeout 300 python -m pytest tests/ -m gpu -q > gpurun_out/f_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/f_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
timeout 100 python bench.py --workload gd_step --steps 5 --warmup 3 > gpurun_out/f_bench_gd.json 2> gpurun_out/f_bench_gd.err
timeout 100 python scripts/timeline_gd_step.py > gpurun_out/f_gd_timeline.txt 2>&1
timeout 60 python bench.py --workload synthesis_fwd --res 256 --steps 10 --warmup 3 > gpurun_out/f_bench_fwd256.json 2> gpurun_out/f_bench_fwd256.err
timeout 60 python bench.py --workload synthesis_fwd --res 1024 --steps 10 --warmup 3 > gpurun_out/f_bench_fwd1024.json 2> gpurun_out/f_bench_fwd1024.err
timeout 100 python bench.py --workload full_loop --steps 2 > gpurun_out/f_bench_full_loop.json 2> gpurun_out/f_bench_full_loop.err
grep -E "^FAILED|passed|failed|exit" gpurun_out/f_tests.log | tail -8; tail -1 gpurun_out/f_smoke.log; for f in f_bench f_bench_gd f_bench_fwd256 f_bench_fwd1024 f_bench_full_loop; do cut -c1-180 gpurun_out/$f.json; tail -1 gpurun_out/$f.err | cut -c1-200; done; head -8 gpurun_out/f_gd_timeline.txt
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/f_ncu_bench.log 2>&1
echo "ncu exit $?"
