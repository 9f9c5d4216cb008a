timeout 120 python tools/probe_mma.py > gpurun_out/r2k_probe_mma.txt 2>&1; head -17 gpurun_out/r2k_probe_mma.txt
timeout 600 python tools/probe_strip.py 10 > gpurun_out/r2k_probe_strip.txt 2>&1; cat gpurun_out/r2k_probe_strip.txt | tail -25
timeout 300 python -m pytest tests/test_gpu_losses.py -q -x 2>&1 | tail -3
timeout 300 python tools/prof_graph.py refine 2>/dev/null > gpurun_out/r2k_graph.txt; grep -A14 "ATen kernels" gpurun_out/r2k_graph.txt
