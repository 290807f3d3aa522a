#!/bin/bash
# round-2 GPU call D: GPU suite, CTA-pair probe (cta_group::2 MMAs) against the cluster-multicast form, gd_step bench + timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest.log
SGV_CONV_PAIR=0 timeout 300 python scripts/pair_probe.py > gpurun_out/d_probe_pair0.jsonl 2> gpurun_out/d_probe_pair0.err
echo "pair0 rc=$?" >> gpurun_out/d_probe_pair0.err
SGV_CONV_PAIR=1 timeout 300 python scripts/pair_probe.py > gpurun_out/d_probe_pair1.jsonl 2> gpurun_out/d_probe_pair1.err
echo "pair1 rc=$?" >> gpurun_out/d_probe_pair1.err
nvidia-smi --query-gpu=name,clocks.sm,memory.used --format=csv > gpurun_out/d_smi_after_probe.txt 2>&1
SGV_CONV_PAIR=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-second-mode > gpurun_out/d_bench_pair1.json 2> gpurun_out/d_bench_pair1.err
echo "bench pair1 rc=$?" >> gpurun_out/d_bench_pair1.err
timeout 600 python bench.py --workload gd_step --steps 5 > gpurun_out/d_bench_gd.json 2> gpurun_out/d_bench_gd.err
echo "gd rc=$?" >> gpurun_out/d_bench_gd.err
timeout 300 python scripts/timeline_gd_step.py > gpurun_out/d_timeline_gd.txt 2>&1
tail -4 gpurun_out/d_pytest.log; cat gpurun_out/d_probe_pair1.err | tail -3
