timeout 300 python -m pytest tests/test_gpu_in_fused.py -q -x 2>&1 | tail -3
for cfg in "102400 16" "102400 8" "65536 16" "49152 16" "36864 16"; do
  set -- $cfg
  echo "=== smem=$1 maxcs=$2"
  EVE_B200_IN_SMEM=$1 EVE_B200_IN_MAXCS=$2 timeout 200 python tools/bench_in.py all 2>&1 | tail -40
done > gpurun_out/r2b_bench_in.txt 2>&1
grep -E "===|total|9216x16 |9216x64|2304x64 |576x128|1024x64|144x256|16x512" gpurun_out/r2b_bench_in.txt
