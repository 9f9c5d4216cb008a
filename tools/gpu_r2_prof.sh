# round-2 profile pass: whole-step launch list with DRAM bytes (eager profiled step of bench.py),
# then ncu --set full captures of the kernels round 2 added / changed, summarised on the box
# (the .ncu-rep files together exceed what gpurun copies back)
start=$(date +%s)
EVE_BENCH_NCU_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
echo "ncu list rc=$? secs=$(( $(date +%s) - start )) lines=$(wc -l < gpurun_out/r2_launches.csv)"
python tools/ncu_launch_summary.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches_summary.txt 2>&1
python tools/ncu_conv_traffic.py gpurun_out/r2_launches.csv 1 gpurun_out/r2_conv_traffic.json > /dev/null 2>&1
gzip -9 gpurun_out/r2_launches.csv
for spec in "strip_fwd:conv_tc_strip_kernel:4" "strip_wgrad:conv_tc_wgrad_strip_kernel:3" "box:conv_tc_kernel:6" "wgrad:conv_tc_wgrad_kernel:3" "cgru:cgru_seq:2" "row:conv_tc_row_kernel:3" "in_bwd:in_bwd_fused_kernel:3" "in_fwd:in_fwd_fused_kernel:2"; do
  name=${spec%%:*}; rest=${spec#*:}; pat=${rest%%:*}; cnt=${rest##*:}
  start=$(date +%s)
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$pat" -c $cnt -o gpurun_out/r2_$name python tools/prof_step.py refine > gpurun_out/r2_ncu_$name.log 2>&1
  python tools/ncu_sum.py gpurun_out/r2_$name.ncu-rep > gpurun_out/r2_ncu_$name.txt 2>&1
  echo "ncu $name rc=$? secs=$(( $(date +%s) - start ))"
  if [ "$name" != "strip_wgrad" ] && [ "$name" != "cgru" ]; then rm -f gpurun_out/r2_$name.ncu-rep; fi
  rm -f gpurun_out/r2_ncu_$name.log
done
du -sh gpurun_out; ls gpurun_out
