for m in 0 2; do
BENCH_IN_STREAM=$m BENCH_IN_SHAPES=240x9216x32 timeout 300 ncu --set full --clock-control none --import-source on -k regex:in_bwd -c 1 -o gpurun_out/r2ae_inbwd_$m python tools/bench_in.py bwd > gpurun_out/r2ae_ncu_$m.log 2>&1; echo "rc=$?"
done
ls -la gpurun_out/
