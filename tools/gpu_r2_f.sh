timeout 300 python tools/conv_table.py refine 3 > gpurun_out/r2f_conv_table.txt 2>&1; head -60 gpurun_out/r2f_conv_table.txt
timeout 600 python tools/grad_precision.py 2 > gpurun_out/r2f_gradprec2.txt 2>&1; tail -8 gpurun_out/r2f_gradprec2.txt
timeout 900 python tools/grad_precision.py 3 > gpurun_out/r2f_gradprec3.txt 2>&1; tail -12 gpurun_out/r2f_gradprec3.txt
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_bench_size.py::test_config3_full_eve_step_matches_the_oracle_at_B8_T30 --deselect tests/test_gpu_bench_size.py::test_config2_static_eyenet_step_matches_the_oracle_at_B8_T30 > gpurun_out/r2f_pytest.txt 2>&1; tail -5 gpurun_out/r2f_pytest.txt
