mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_upfirdn2d_gpu.py tests/test_synthesis_gpu.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench_r1ac.err | tee gpurun_out/bench_r1ac.json | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1ac.csv python scripts/profile_step.py > gpurun_out/profile_step_r1ac.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_r1ac.csv 40
