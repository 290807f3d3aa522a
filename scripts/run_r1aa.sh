mkdir -p gpurun_out
timeout 120 scripts/_bin/mma_rate_probe 2>&1 | tee gpurun_out/mma_rate_probe_r1.txt
for dbg in 0 1 2 3 4 5 6 7; do
  echo "== v3 debug=$dbg"; SGV_CONV_CLUSTER=1 SGV_V3_DEBUG=$dbg timeout 200 python scripts/bench_conv.py main4 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['kernel'], round(d['ms'],3), round(d['tflops']))"
done
