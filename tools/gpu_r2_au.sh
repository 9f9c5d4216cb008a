timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_ops.py -x -q -m gpu > gpurun_out/r2au_pytest.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2au_pytest.txt
timeout 600 python tools/conv_table.py > gpurun_out/r2au_conv_table.txt 2>&1; head -1 gpurun_out/r2au_conv_table.txt
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2au_bench.json 2> gpurun_out/r2au_bench.err; echo "rc=$?"; cut -c1-230 gpurun_out/r2au_bench.json
