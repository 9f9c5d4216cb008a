start=$(date +%s)
timeout 300 python -m pytest tests/ -q -x -m gpu > gpurun_out/pytest_final.txt 2>&1
echo "pytest rc=$? secs=$(( $(date +%s) - start ))"; tail -3 gpurun_out/pytest_final.txt
start=$(date +%s)
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$? secs=$(( $(date +%s) - start ))"
start=$(date +%s)
EVE_BENCH_NCU_RANGE=1 timeout 330 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/launches_final_bench.log 2>&1
echo "ncu list rc=$? secs=$(( $(date +%s) - start )) lines=$(wc -l < gpurun_out/launches_final.csv)"
start=$(date +%s)
timeout 150 ncu --set full --clock-control none -k regex:'conv_tc_row_kernel|conv_tc_wgrad_row_kernel' -c 10 -o gpurun_out/tc_row python tools/prof_step.py refine > gpurun_out/ncu_row.log 2>&1
echo "ncu row rc=$? secs=$(( $(date +%s) - start ))"
ls -la gpurun_out | tail -8
