timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2p_pytest.txt 2>&1; tail -6 gpurun_out/r2p_pytest.txt
timeout 300 python tools/prof_graph.py refine 2>/dev/null > gpurun_out/r2p_graph.txt; head -50 gpurun_out/r2p_graph.txt
