timeout 900 python -m pytest tests/test_gpu_in_fused.py -x -q -m gpu > gpurun_out/r2aj_pytest.txt 2>&1; tail -3 gpurun_out/r2aj_pytest.txt
BENCH_IN_STREAM=2 timeout 300 python tools/bench_in.py bwd > gpurun_out/r2aj_in_bwd_ring.txt 2>&1
paste <(cut -c1-42 gpurun_out/r2ad_in_bwd_staged.txt) <(cut -c19-42 gpurun_out/r2ad_in_bwd_stream.txt) <(cut -c19-42 gpurun_out/r2aj_in_bwd_ring.txt)
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err; cut -c1-230 gpurun_out/r2aj_bench.json
