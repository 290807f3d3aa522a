cd scripts
echo "== per-tap kernel"; SGV_CONV_NO_V3=1 timeout 120 python bench_dgrad_up.py
echo "== persistent kernel"; timeout 120 python bench_dgrad_up.py
echo "== persistent, debug=2 (no epilogue)"; SGV_V3_DEBUG=2 timeout 120 python bench_dgrad_up.py
echo "== persistent, debug=6 (loads+transform only)"; SGV_V3_DEBUG=6 timeout 120 python bench_dgrad_up.py
