export PROBE_STRIP_GEOMS=480x32x32x64x64
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 1 -o gpurun_out/r2ax_box python tools/probe_strip.py 2 > gpurun_out/r2ax_ncu.log 2>&1; echo "rc=$?"
