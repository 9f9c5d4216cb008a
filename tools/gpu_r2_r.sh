timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2r_pytest.txt 2>&1; tail -4 gpurun_out/r2r_pytest.txt
timeout 300 python tools/conv_table.py refine 3 > gpurun_out/r2r_conv_table.txt 2>&1; head -12 gpurun_out/r2r_conv_table.txt
timeout 300 python tools/prof_graph.py refine 2>/dev/null > gpurun_out/r2r_graph.txt; head -3 gpurun_out/r2r_graph.txt
timeout 1500 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2r_bench.json
