#!/bin/bash
# round-2 GPU call B: full GPU suite, smoke, bench (both arms), launch list of the step, one ncu --set full capture of the hot kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=10 > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/b_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/b_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
echo "bench rc=$?" >> gpurun_out/b_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b_bench_ref.json 2> gpurun_out/b_bench_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_tf32_v3|wgrad_tf32_v2|wgrad_tf32_s64|fir_nhwc_tma44' -o gpurun_out/ncu_r2b -f python scripts/ncu_r2_target.py > gpurun_out/b_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/b_ncu.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-second-mode > gpurun_out/b_launch_bench.log 2>&1
echo "launches rc=$?" >> gpurun_out/b_launch_bench.log
tail -5 gpurun_out/b_pytest.log
