mkdir -p gpurun_out
timeout 600 python scripts/timeline_step.py > gpurun_out/timeline_r1al.txt 2>gpurun_out/timeline_r1al.err; tail -5 gpurun_out/timeline_r1al.err; cat gpurun_out/timeline_r1al.txt
