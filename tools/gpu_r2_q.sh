timeout 900 python -m pytest tests/test_gpu_options.py -q -x -k "strip_weight or wave" > gpurun_out/r2q_pytest1.txt 2>&1; tail -8 gpurun_out/r2q_pytest1.txt
timeout 300 python tools/conv_table.py refine 3 > gpurun_out/r2q_conv_table.txt 2>&1; head -40 gpurun_out/r2q_conv_table.txt
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2q_pytest.txt 2>&1; tail -6 gpurun_out/r2q_pytest.txt
