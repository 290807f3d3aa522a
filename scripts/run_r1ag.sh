cd scripts
for mb in 256 128; do
echo "== MAXBN=$mb"; SGV_V3_MAXBN=$mb timeout 120 python bench_dgrad_up.py; SGV_V3_MAXBN=$mb timeout 200 python bench_conv.py main4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print(d['kernel'], round(d['ms'],3), round(d['tflops']))
    except Exception: print(l.rstrip()[:200])"
done
