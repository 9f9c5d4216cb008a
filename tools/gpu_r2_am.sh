timeout 900 python -m pytest tests/test_gpu_in_fused.py tests/test_gpu_models.py -x -q -m gpu > gpurun_out/r2am_pytest.txt 2>&1; tail -3 gpurun_out/r2am_pytest.txt
BENCH_IN_STREAM=2 timeout 300 python tools/bench_in.py bwd > gpurun_out/r2am_in_bwd.txt 2>&1; tail -1 gpurun_out/r2am_in_bwd.txt
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2am_bench.json 2> gpurun_out/r2am_bench.err; cut -c1-230 gpurun_out/r2am_bench.json
EVE_B200_IN_STREAM_CS=1 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2am_bench_cs1.json 2> gpurun_out/r2am_bench_cs1.err; cut -c1-230 gpurun_out/r2am_bench_cs1.json
