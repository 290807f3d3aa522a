mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_synthesis_gpu.py -x -q 2>&1 | tail -3
for aw in 0 1; do
echo "== async_wgrad=$aw"; SGV_ASYNC_WGRAD=$aw timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench_r1ak_$aw.err | tee gpurun_out/bench_r1ak_$aw.json | cut -c1-330; tail -2 gpurun_out/bench_r1ak_$aw.err
done
