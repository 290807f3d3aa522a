mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_upfirdn2d_gpu.py tests/test_synthesis_gpu.py -x -q 2>&1 | tail -3
timeout 200 python scripts/bench_ops.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['kernel'], round(d['ms'],3), round(d['gbs']))"
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench_r1aj.err | tee gpurun_out/bench_r1aj.json | cut -c1-330
