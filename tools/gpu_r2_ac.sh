# wgrad-row single-lane loop check + fresh graph timeline
timeout 900 python -m pytest tests/test_gpu_options.py tests/test_gpu_ops.py -x -q -m gpu > gpurun_out/r2ac_pytest.txt 2>&1; tail -3 gpurun_out/r2ac_pytest.txt
timeout 600 python tools/conv_table.py > gpurun_out/r2ac_conv_table.txt 2>&1; head -1 gpurun_out/r2ac_conv_table.txt
timeout 600 python tools/prof_graph.py refine > gpurun_out/r2ac_graph_step.txt 2>&1; head -50 gpurun_out/r2ac_graph_step.txt
