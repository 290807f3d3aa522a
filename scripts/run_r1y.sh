mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_synthesis_gpu.py -x -q 2>&1 | tail -2
SGV_CONV_CLUSTER=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tf32_v3 --launch-skip 1 --launch-count 5 -o gpurun_out/v3_r1y -f python scripts/ncu_v3_target.py > gpurun_out/ncu_v3_r1y.log 2>&1
tail -3 gpurun_out/ncu_v3_r1y.log
ls -la gpurun_out/*.ncu-rep
