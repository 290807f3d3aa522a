mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py 2>gpurun_out/bench_final.err | tee gpurun_out/bench_final.json | cut -c1-400
tail -3 gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json | cut -c1-400
