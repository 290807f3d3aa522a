mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_synthesis_gpu.py tests/test_abi.py -x -q 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench_r1ad.err | tee gpurun_out/bench_r1ad.json | cut -c1-330
