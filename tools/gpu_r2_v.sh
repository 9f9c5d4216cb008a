timeout 300 python tools/prof_graph.py refine 2>/dev/null > gpurun_out/r2v_graph.txt; head -3 gpurun_out/r2v_graph.txt
timeout 1500 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2v_bench.json
timeout 120 python tools/probe_mma.py > gpurun_out/r2v_probe_mma.txt 2>&1; tail -14 gpurun_out/r2v_probe_mma.txt
