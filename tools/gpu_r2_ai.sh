timeout 600 python tools/prof_graph.py refine > gpurun_out/r2ai_graph_step.txt 2>&1; grep -E "stem|kernels |conv_tc_kernel<64|in_stats|in_relu_maxpool" gpurun_out/r2ai_graph_step.txt
