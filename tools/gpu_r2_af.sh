timeout 900 python -m pytest tests/test_gpu_in_fused.py tests/test_gpu_options.py -x -q -m gpu > gpurun_out/r2af_pytest.txt 2>&1; tail -3 gpurun_out/r2af_pytest.txt
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err; cut -c1-330 gpurun_out/r2af_bench.json
EVE_B200_IN_STREAM=0 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2af_bench_staged.json 2> gpurun_out/r2af_bench_staged.err; cut -c1-330 gpurun_out/r2af_bench_staged.json
EVE_B200_IN_STREAM=2 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2af_bench_always.json 2> gpurun_out/r2af_bench_always.err; cut -c1-330 gpurun_out/r2af_bench_always.json
