#!/bin/bash
# round-2 closing call (1 GPU): smoke, full GPU suite, every bench workload + the reference arm, ncu launch list of the step, ncu --set full
# capture of the hot kernels (-> profiles/ncu_traffic.json via scripts/ncu_traffic.py, run on the build box)
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/z_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/z_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/z_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z_pytest.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err
echo "bench rc=$?" >> gpurun_out/z_bench.err
timeout 600 python bench.py --workload gd_step --steps 5 > gpurun_out/z_bench_gd.json 2> gpurun_out/z_bench_gd.err
timeout 300 python bench.py --workload synthesis_fwd --res 256 --steps 10 > gpurun_out/z_bench_fwd256.json 2> gpurun_out/z_bench_fwd256.err
timeout 300 python bench.py --workload synthesis_fwd --res 1024 --steps 10 > gpurun_out/z_bench_fwd1024.json 2> gpurun_out/z_bench_fwd1024.err
timeout 900 python bench.py --workload full_loop --steps 2 > gpurun_out/z_bench_full_loop.json 2> gpurun_out/z_bench_full_loop.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_tf32_v3|wgrad_tf32_v2|wgrad_tf32_s64|fir_nhwc_tma44' -o gpurun_out/ncu_r2z -f python scripts/ncu_r2_target.py > gpurun_out/z_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/z_ncu.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-second-mode > gpurun_out/z_launch_bench.log 2>&1
echo "launches rc=$?" >> gpurun_out/z_launch_bench.log
tail -4 gpurun_out/z_pytest.log; tail -2 gpurun_out/z_smoke.log; for f in z_bench z_bench_gd z_bench_fwd256 z_bench_fwd1024 z_bench_full_loop z_bench_ref; do head -c 220 gpurun_out/$f.json; echo; done
