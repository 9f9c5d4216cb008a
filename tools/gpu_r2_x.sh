timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2x_pytest.txt 2>&1; tail -4 gpurun_out/r2x_pytest.txt
timeout 300 python tools/prof_graph.py refine 2>/dev/null > gpurun_out/r2x_graph.txt; head -3 gpurun_out/r2x_graph.txt; grep compact gpurun_out/r2x_graph.txt
