#!/bin/bash
# round-2 GPU call F: GPU suite, CUPTI timeline of the graph step, ncu --set full capture of the hot kernels (traffic json)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest.log
timeout 300 python scripts/timeline_step.py > gpurun_out/f_timeline.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_tf32_v3|wgrad_tf32_v2|wgrad_tf32_s64|fir_nhwc_tma44' -o gpurun_out/ncu_r2f -f python scripts/ncu_r2_target.py > gpurun_out/f_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/f_ncu.log
tail -4 gpurun_out/f_pytest.log
