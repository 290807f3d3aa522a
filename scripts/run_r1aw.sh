mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fir_nhwc_tma44|conv_tf32_v3|wgrad_tf32_v2|modconv_act_bwd" --launch-skip 10 --launch-count 10 -o gpurun_out/final_r1aw -f python scripts/ncu_final_target.py > gpurun_out/ncu_final_r1aw.log 2>&1
tail -2 gpurun_out/ncu_final_r1aw.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1aw.csv python scripts/profile_step.py > gpurun_out/profile_step_r1aw.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_r1aw.csv 30
