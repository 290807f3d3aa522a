mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench_r1am.err | tee gpurun_out/bench_r1am.json | cut -c1-330
