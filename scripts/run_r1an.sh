mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_synthesis_gpu.py -x -q 2>&1 | tail -3
cd scripts
for dbg in 0 2; do
echo "== v3 debug=$dbg"; SGV_V3_DEBUG=$dbg timeout 200 python bench_conv.py main4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print(d['kernel'], 'conv', round(d['ms'],3), round(d['tflops']), 'wgrad', round(d['wgrad_ms'],3), round(d['wgrad_tflops']))
    except Exception: print(l.rstrip()[:200])"
done
cd ..
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench_r1an.err | tee gpurun_out/bench_r1an.json | cut -c1-330
