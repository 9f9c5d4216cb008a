timeout 300 python -m pytest tests/test_gpu_in_fused.py -q -x 2>&1 | tail -3
for cfg in "102400 8" "65536 8" "49152 8"; do
  set -- $cfg
  echo "=== smem=$1 maxcs=$2"
  EVE_B200_IN_SMEM=$1 EVE_B200_IN_MAXCS=$2 timeout 200 python tools/bench_in.py all 2>&1 | tail -40
done > gpurun_out/r2e_bench_in.txt 2>&1
grep -E "===|total|9216x16 |9216x64|2304x64 |576x128|1024x64|144x256|16x512" gpurun_out/r2e_bench_in.txt
timeout 200 python tools/prof_graph.py refine 2>/dev/null > gpurun_out/r2e_graph_fused.txt; head -12 gpurun_out/r2e_graph_fused.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; cut -c1-330 gpurun_out/r2e_bench.json
