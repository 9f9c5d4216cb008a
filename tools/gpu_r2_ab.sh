timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_row_kernel -c 1 -o gpurun_out/r2ab_row python tools/prof_step.py refine > gpurun_out/r2ab_ncu.log 2>&1; echo "rc=$?"
ls -la gpurun_out/
