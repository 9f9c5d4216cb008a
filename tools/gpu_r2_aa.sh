# row-kernel rework check: conv parity tests, per-layer table, short bench
timeout 900 python -m pytest tests/test_gpu_options.py tests/test_gpu_ops.py tests/test_gpu_graph.py -x -q -m gpu > gpurun_out/r2aa_pytest.txt 2>&1; tail -3 gpurun_out/r2aa_pytest.txt
timeout 600 python tools/conv_table.py > gpurun_out/r2aa_conv_table.txt 2>&1; grep -i "row" gpurun_out/r2aa_conv_table.txt | head -40
timeout 600 python bench.py --steps 40 --warmup 5 > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; cat gpurun_out/r2aa_bench.json | cut -c1-600
