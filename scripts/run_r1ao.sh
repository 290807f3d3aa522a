mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_synthesis_gpu.py tests/test_dropin_gpu.py -x -q 2>&1 | tail -3
cd scripts
for m in 0 1; do
echo "== MH4=$m"; SGV_V3_MH4=$m timeout 200 python bench_conv.py b256.conv1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print(d['kernel'], 'conv', round(d['ms'],3), round(d['tflops']), 'wgrad', round(d['wgrad_ms'],3), round(d['wgrad_tflops']))
    except Exception: print(l.rstrip()[:200])"
SGV_V3_MH4=$m timeout 100 python bench_dgrad_up.py | head -1
done
cd ..
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench_r1ao.err | tee gpurun_out/bench_r1ao.json | cut -c1-330
