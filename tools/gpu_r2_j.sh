timeout 900 python -m pytest tests/test_gpu_losses.py tests/test_gpu_gaze.py -q -x > gpurun_out/r2j_pytest1.txt 2>&1; tail -15 gpurun_out/r2j_pytest1.txt
timeout 1200 python -m pytest tests/test_gpu_bench_size.py -q -x -s -k "config" > gpurun_out/r2j_pytest2.txt 2>&1; tail -4 gpurun_out/r2j_pytest2.txt; grep -E "gradients vs fp64|forward rel" gpurun_out/r2j_pytest2.txt | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_bench_size.py > gpurun_out/r2j_pytest3.txt 2>&1; tail -12 gpurun_out/r2j_pytest3.txt
timeout 300 python tools/prof_graph.py refine 2>/dev/null > gpurun_out/r2j_graph.txt; head -4 gpurun_out/r2j_graph.txt; grep -c "at::native" gpurun_out/r2j_graph.txt
