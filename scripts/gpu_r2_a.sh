#!/bin/bash
# round-2 GPU call A: full GPU suite (no -x: collect every failure), smoke, bench (both arms)
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/a_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?" >> gpurun_out/a_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err
echo "ref rc=$?" >> gpurun_out/a_bench_ref.err
tail -5 gpurun_out/a_pytest.log
