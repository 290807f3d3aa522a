mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_synthesis_gpu.py tests/test_dropin_gpu.py -x -q 2>&1 | tail -3
for dbg in 0 2; do
  echo "== v3 debug=$dbg"; SGV_V3_DEBUG=$dbg timeout 200 python scripts/bench_conv.py main4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print(d['kernel'], round(d['ms'],3), round(d['tflops']))
    except Exception: print(l.rstrip()[:200])"
done
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench_r1ab.err | tee gpurun_out/bench_r1ab.json | cut -c1-330
