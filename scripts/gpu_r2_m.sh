#!/bin/bash
# round-2 GPU call M: reference loss phases on the drop-in ops (new test), serial-time / tail analysis of the headline step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_reference_on_dropin_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/m_pytest.log
timeout 300 python scripts/timeline_step.py > gpurun_out/m_timeline.txt 2>&1
tail -12 gpurun_out/m_pytest.log | cut -c1-600
