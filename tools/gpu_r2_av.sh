timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_ops.py tests/test_gpu_options.py -x -q -m gpu > gpurun_out/r2av_pytest.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2av_pytest.txt
timeout 600 python tools/conv_table.py > gpurun_out/r2av_conv_table.txt 2>&1; head -1 gpurun_out/r2av_conv_table.txt
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2av_bench.json 2> gpurun_out/r2av_bench.err; echo "rc=$?"; cut -c1-230 gpurun_out/r2av_bench.json
