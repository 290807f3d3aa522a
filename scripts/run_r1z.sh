mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_synthesis_gpu.py -x -q 2>&1 | tail -2
for cl in 1 2; do
  echo "== bench_conv cluster=$cl"; SGV_CONV_CLUSTER=$cl timeout 200 python scripts/bench_conv.py 2>gpurun_out/bench_conv_r1z_$cl.err | tee gpurun_out/bench_conv_r1z_$cl.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['kernel'], round(d['ms'],3), round(d['tflops']), 'wgrad', round(d['wgrad_ms'],3), round(d['wgrad_tflops']))"
done
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench_r1z.err | tee gpurun_out/bench_r1z.json | cut -c1-330
