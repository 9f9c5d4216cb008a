timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_ops.py tests/test_gpu_options.py -x -q -m gpu > gpurun_out/r2ay_pytest.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2ay_pytest.txt
for m in 1 0; do
EVE_B200_TC_DUAL=$m timeout 600 python tools/conv_table.py > gpurun_out/r2ay_conv_table_$m.txt 2>&1; head -1 gpurun_out/r2ay_conv_table_$m.txt
EVE_B200_TC_DUAL=$m timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2ay_bench_$m.json 2> gpurun_out/r2ay_bench_$m.err; echo "rc=$?"; cut -c1-230 gpurun_out/r2ay_bench_$m.json
done
