mkdir -p gpurun_out
for cfg in "512 2" "256 2" "512 1"; do set -- $cfg
echo "== WIDE_CIN=$1 CLUSTER=$2"; SGV_V3_WIDE_CIN=$1 SGV_CONV_CLUSTER=$2 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-200
done
