mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_synthesis_gpu.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench_r1ar.err | tee gpurun_out/bench_r1ar.json | cut -c1-200
