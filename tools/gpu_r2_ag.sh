timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2ag_pytest.txt 2>&1; tail -5 gpurun_out/r2ag_pytest.txt
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2ag_bench.json 2> gpurun_out/r2ag_bench.err; cut -c1-330 gpurun_out/r2ag_bench.json
