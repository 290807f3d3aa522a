cd scripts
for dbg in 0 1 2 4 6 7; do
echo "== wgrad debug=$dbg"; SGV_WG_DEBUG=$dbg timeout 200 python bench_conv.py main4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print(d['kernel'], 'conv', round(d['ms'],3), round(d['tflops']), 'wgrad', round(d['wgrad_ms'],3), round(d['wgrad_tflops']))
    except Exception: print(l.rstrip()[:200])"
done
