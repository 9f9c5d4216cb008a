# round-2 final pass on one B200: tests, whole-step ncu launch list (+ DRAM bytes), ncu --set full
# summaries of the hot kernels, per-layer table, graph timeline, then the bench line itself
start=$(date +%s)
timeout 900 python -m pytest tests/ -q -x -m gpu > gpurun_out/r2f_pytest.txt 2>&1
echo "pytest rc=$? secs=$(( $(date +%s) - start ))"; tail -3 gpurun_out/r2f_pytest.txt
start=$(date +%s)
EVE_BENCH_NCU_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 1 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/r2f_launches_bench.log 2>&1
echo "ncu list rc=$? secs=$(( $(date +%s) - start )) lines=$(wc -l < gpurun_out/r2f_launches.csv)"
python tools/ncu_launch_summary.py gpurun_out/r2f_launches.csv > gpurun_out/r2f_launches_summary.txt 2>&1
python tools/ncu_conv_traffic.py gpurun_out/r2f_launches.csv 1 gpurun_out/r2f_conv_traffic.json > /dev/null 2>&1
cp gpurun_out/r2f_conv_traffic.json profiles/conv_traffic.json
gzip -9 gpurun_out/r2f_launches.csv
for spec in "strip_fwd:conv_tc_strip_kernel:4" "strip_wgrad:conv_tc_wgrad_strip_kernel:3" "box:conv_tc_kernel:6" "wgrad:conv_tc_wgrad_kernel:3" "cgru:cgru_seq:2" "row:conv_tc_row_kernel:4" "wgrad_row:conv_tc_wgrad_row_kernel:3" "in_bwd:in_bwd_stream_kernel:4" "in_fwd:in_fwd_fused_kernel:2"; do
  name=${spec%%:*}; rest=${spec#*:}; pat=${rest%%:*}; cnt=${rest##*:}
  start=$(date +%s)
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$pat" -c $cnt -o gpurun_out/r2f_$name python tools/prof_step.py refine > gpurun_out/r2f_ncu_$name.log 2>&1
  python tools/ncu_sum.py gpurun_out/r2f_$name.ncu-rep > gpurun_out/r2f_ncu_$name.txt 2>&1
  echo "ncu $name rc=$? secs=$(( $(date +%s) - start ))"
  rm -f gpurun_out/r2f_$name.ncu-rep gpurun_out/r2f_ncu_$name.log
done
timeout 600 python tools/conv_table.py > gpurun_out/r2f_conv_table.txt 2>&1; head -1 gpurun_out/r2f_conv_table.txt
timeout 600 python tools/prof_graph.py refine > gpurun_out/r2f_graph_step.txt 2>&1; grep "kernels " gpurun_out/r2f_graph_step.txt
start=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$? secs=$(( $(date +%s) - start ))"; cut -c1-300 gpurun_out/r2f_bench.json
du -sh gpurun_out
