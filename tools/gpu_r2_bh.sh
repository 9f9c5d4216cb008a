timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py tests/test_gpu_bench_size.py -q -x -m gpu > gpurun_out/r2bh_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2bh_pytest.txt
timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extra --no-e2e > gpurun_out/r2bh_bench.json 2> gpurun_out/r2bh_bench.err; echo "bench rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2bh_bench.json')); print(j['ms_per_step'], j['value'], j['roofline']['frac'], j['clocks'])"
timeout 300 python tools/prof_graph.py refine > gpurun_out/r2bh_graph_step.txt 2>&1; grep "kernels \|im2col\|in_stats" gpurun_out/r2bh_graph_step.txt
