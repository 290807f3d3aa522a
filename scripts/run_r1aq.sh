mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -m gpu -x -q 2>&1 | tail -3
for ps in 0 1; do
echo "== prescale=$ps"; SGV_PRESCALE=$ps timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench_r1aq_$ps.err | tee gpurun_out/bench_r1aq_$ps.json | cut -c1-200
done
