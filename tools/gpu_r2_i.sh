for m in 7 1 3 5; do
  echo "=== TC_MASK=$m"; EVE_B200_TC_MASK=$m timeout 600 python tools/grad_precision.py 3 4 15 2>&1 | grep -E "^(eye_net|refine_net):|g_final|oracle"
done
echo "=== TC_MASK=7 FUSED_NORM=0"; EVE_B200_FUSED_NORM=0 timeout 600 python tools/grad_precision.py 3 4 15 2>&1 | grep -E "^(eye_net|refine_net):|g_final"
echo "=== TC_MASK=7 FUSED_PLANES=0"; EVE_B200_FUSED_PLANES=0 timeout 600 python tools/grad_precision.py 3 4 15 2>&1 | grep -E "^(eye_net|refine_net):|g_final"
timeout 1500 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2i_bench.json; tail -3 gpurun_out/r2i_bench.err
