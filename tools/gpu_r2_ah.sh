timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "stem or conv2d" > gpurun_out/r2ah_pytest.txt 2>&1; tail -15 gpurun_out/r2ah_pytest.txt
for m in 0 1 2; do
EVE_B200_STEM_WINDOWS=$m timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2ah_bench_$m.json 2> gpurun_out/r2ah_bench_$m.err; cut -c1-230 gpurun_out/r2ah_bench_$m.json
done
