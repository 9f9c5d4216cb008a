timeout 900 python -m pytest tests/test_gpu_options.py tests/test_gpu_ops.py -q -x > gpurun_out/r2u_pytest1.txt 2>&1; tail -8 gpurun_out/r2u_pytest1.txt
timeout 300 python tools/conv_table.py refine 3 > gpurun_out/r2u_conv_table.txt 2>&1; head -3 gpurun_out/r2u_conv_table.txt; grep wgrad gpurun_out/r2u_conv_table.txt | head -24
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2u_pytest.txt 2>&1; tail -4 gpurun_out/r2u_pytest.txt
