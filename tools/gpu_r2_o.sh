timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2o_pytest.txt 2>&1; tail -6 gpurun_out/r2o_pytest.txt
timeout 300 python tools/prof_graph.py refine 2>/dev/null > gpurun_out/r2o_graph.txt; head -36 gpurun_out/r2o_graph.txt
timeout 600 python bench.py --leg stream900 2>&1 | tail -1 | cut -c1-200
