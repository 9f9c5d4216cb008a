timeout 900 python -m pytest tests/test_gpu_bench_size.py -q -x -s -k "config" > gpurun_out/r2h_pytest.txt 2>&1; tail -4 gpurun_out/r2h_pytest.txt; grep -E "gradients vs fp64|forward rel" gpurun_out/r2h_pytest.txt | cut -c1-400
start=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$? secs=$(( $(date +%s) - start ))"; cut -c1-400 gpurun_out/r2h_bench.json; tail -3 gpurun_out/r2h_bench.err
start=$(date +%s)
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2h_ref.json 2> gpurun_out/r2h_ref.err; echo "ref rc=$? secs=$(( $(date +%s) - start ))"; cut -c1-300 gpurun_out/r2h_ref.json
