#!/bin/bash
# round-2 GPU call S: dense data-gradient / weight-gradient kernels restructured (8 weight rows in flight; 16 weight rows per CTA for long
# reductions): dense + synthesis + train-step tests, then the headline line
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_dense_gpu.py tests/test_synthesis_gpu.py tests/test_train_aux_gpu.py tests/test_networks_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-second-mode --no-cpu-baseline > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/s_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])
PY
