timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_graph.py -x -q -m gpu > gpurun_out/r2aw_pytest.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2aw_pytest.txt
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2aw_bench.json 2> gpurun_out/r2aw_bench.err; echo "rc=$?"; cut -c1-230 gpurun_out/r2aw_bench.json
timeout 600 python tools/prof_graph.py refine > gpurun_out/r2aw_graph_step.txt 2>&1; sed -n 5,45p gpurun_out/r2aw_graph_step.txt | cut -c1-110
