timeout 200 python tools/prof_graph.py refine 2>/dev/null > gpurun_out/r2d_graph_fused.txt; echo rc=$?
EVE_B200_FUSED_NORM=0 timeout 200 python tools/prof_graph.py refine 2>/dev/null > gpurun_out/r2d_graph_legacy.txt; echo rc=$?
cat gpurun_out/r2d_graph_fused.txt; head -40 gpurun_out/r2d_graph_legacy.txt
start=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_bench_size.py tests/test_gpu_graph.py tests/test_gpu_options.py -q > gpurun_out/r2d_pytest.txt 2>&1
echo "pytest rc=$? secs=$(( $(date +%s) - start ))"; tail -8 gpurun_out/r2d_pytest.txt; grep -E "config[23] (forward|gradient)" gpurun_out/r2d_pytest.txt
