# last pass of round 2: all GPU tests + smoke, per-layer table, graph timeline, the bench line
start=$(date +%s)
timeout 900 python -m pytest tests/ -q -m gpu > gpurun_out/r2h_pytest.txt 2>&1
echo "pytest rc=$? secs=$(( $(date +%s) - start ))"; tail -3 gpurun_out/r2h_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python tools/conv_table.py > gpurun_out/r2h_conv_table.txt 2>&1; head -1 gpurun_out/r2h_conv_table.txt
timeout 300 python tools/prof_graph.py refine > gpurun_out/r2h_graph_step.txt 2>&1; grep "kernels " gpurun_out/r2h_graph_step.txt
start=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench rc=$? secs=$(( $(date +%s) - start ))"; cut -c1-300 gpurun_out/r2h_bench.json
