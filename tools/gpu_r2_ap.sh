timeout 300 python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu -k "pairs" > gpurun_out/r2ap_pytest.txt 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2ap_pytest.txt
for m in 0 1; do
EVE_B200_TC_PAIR=$m timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2ap_bench_$m.json 2> gpurun_out/r2ap_bench_$m.err; echo "rc=$?"; cut -c1-230 gpurun_out/r2ap_bench_$m.json
done
