timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_options.py tests/test_gpu_ops.py -q -x > gpurun_out/r2l_pytest1.txt 2>&1; tail -8 gpurun_out/r2l_pytest1.txt
timeout 600 python tools/probe_strip.py 10 > gpurun_out/r2l_probe_strip.txt 2>&1; tail -22 gpurun_out/r2l_probe_strip.txt
timeout 300 python tools/conv_table.py refine 3 > gpurun_out/r2l_conv_table.txt 2>&1; head -45 gpurun_out/r2l_conv_table.txt
timeout 300 python -m pytest tests/test_gpu_input.py -q -x 2>&1 | tail -5
