mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_synthesis_gpu.py tests/test_dropin_gpu.py -x -q 2>&1 | tail -3
cd scripts; timeout 200 python bench_conv.py main4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print(d['kernel'], 'conv', round(d['ms'],3), round(d['tflops']), 'wgrad', round(d['wgrad_ms'],3), round(d['wgrad_tflops']))
    except Exception: print(l.rstrip()[:200])"; cd ..
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench_r1au.err | tee gpurun_out/bench_r1au.json | cut -c1-200
