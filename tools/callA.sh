set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python tools/prof_step.py refine > gpurun_out/prof_step_refine.txt 2>&1
tail -60 gpurun_out/prof_step_refine.txt
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 1 --warmup 1 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
wc -l gpurun_out/launches_tc.csv
timeout 300 ncu --set full --clock-control none -k regex:conv_tc_kernel -c 16 -o gpurun_out/tc_fwd python tools/prof_step.py refine > gpurun_out/ncu_fwd.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:conv_tc_wgrad -c 10 -o gpurun_out/tc_wgrad python tools/prof_step.py refine > gpurun_out/ncu_wgrad.log 2>&1
ls -la gpurun_out
