export PROBE_STRIP_GEOMS=480x32x32x64x64
EVE_B200_STRIP_R=7 EVE_B200_STRIP_KC=64 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_strip_kernel -c 1 -o gpurun_out/r2as_strip python tools/probe_strip.py 2 > gpurun_out/r2as_ncu.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
