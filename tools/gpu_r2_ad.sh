# streaming InstanceNorm backward: parity, then per-shape timing staged (0) vs streaming (2)
timeout 900 python -m pytest tests/test_gpu_in_fused.py -x -q -m gpu > gpurun_out/r2ad_pytest.txt 2>&1; tail -5 gpurun_out/r2ad_pytest.txt
BENCH_IN_STREAM=0 timeout 300 python tools/bench_in.py bwd > gpurun_out/r2ad_in_bwd_staged.txt 2>&1
BENCH_IN_STREAM=2 timeout 300 python tools/bench_in.py bwd > gpurun_out/r2ad_in_bwd_stream.txt 2>&1
paste <(cut -c1-42 gpurun_out/r2ad_in_bwd_staged.txt) <(cut -c19-42 gpurun_out/r2ad_in_bwd_stream.txt)
