cd scripts
for dbg in 0 1 2 4 6 7; do
echo "== v3 MH4 debug=$dbg"; SGV_V3_DEBUG=$dbg timeout 200 python bench_conv.py b256.conv1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print(d['kernel'], 'conv', round(d['ms'],3), round(d['tflops']))
    except Exception: print(l.rstrip()[:200])"
done
