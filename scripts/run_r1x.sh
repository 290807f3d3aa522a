mkdir -p gpurun_out
for cl in 2 4; do
  echo "== tests cluster=$cl"; SGV_CONV_CLUSTER=$cl timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_synthesis_gpu.py -x -q 2>&1 | tail -3
done
for cl in 1 2 4; do
  echo "== bench_conv cluster=$cl"; SGV_CONV_CLUSTER=$cl timeout 200 python scripts/bench_conv.py 2>gpurun_out/bench_conv_r1x_$cl.err | tee gpurun_out/bench_conv_r1x_$cl.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: (round(v,3) if isinstance(v,float) else v) for k,v in d.items()})"
done
for ps in 0 1; do
  echo "== bench param_stream=$ps"; SGV_PARAM_STREAM=$ps timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench_r1x_ps$ps.err | tee gpurun_out/bench_r1x_ps$ps.json | cut -c1-330
done
