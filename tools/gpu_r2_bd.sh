timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_options.py tests/test_gpu_models.py -q -x -m gpu -k "stem_norm or stem_fused or eyenet_cnn_gradients or maxpool" > gpurun_out/r2bd_pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2bd_pytest.txt
for v in 1 0; do
EVE_B200_STEM_FUSED_BWD=$v timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extra --no-e2e > gpurun_out/r2bd_bench_$v.json 2> gpurun_out/r2bd_bench_$v.err; echo "bench fused=$v rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2bd_bench_$v.json')); print(j['ms_per_step'], j['value'], j['roofline']['frac'], j['final_loss'])"
done
