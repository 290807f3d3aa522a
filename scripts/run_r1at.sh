mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench_r1at.err | tee gpurun_out/bench_r1at.json | cut -c1-200
