for v in 2 1; do
EVE_B200_TC_ROW_WGRAD=$v timeout 200 python tools/conv_table.py > gpurun_out/r2bf_conv_table_$v.txt 2>&1; head -1 gpurun_out/r2bf_conv_table_$v.txt; grep "^wgrad  *240  *72  *128" gpurun_out/r2bf_conv_table_$v.txt
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_wgrad_row3_kernel" -c 3 -o gpurun_out/r2bf_row3 python tools/prof_step.py refine > gpurun_out/r2bf_ncu_row3.log 2>&1
python tools/ncu_sum.py gpurun_out/r2bf_row3.ncu-rep > gpurun_out/r2bf_ncu_row3.txt 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2bf*
