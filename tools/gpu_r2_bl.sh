timeout 400 python -m pytest tests/test_gpu_models.py tests/test_gpu_ops.py -q -x -m gpu > gpurun_out/r2bl_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2bl_pytest.txt
timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-extra --no-e2e > gpurun_out/r2bl_bench.json 2> gpurun_out/r2bl_bench.err; echo "bench rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2bl_bench.json')); print(j['ms_per_step'], j['value'], j['roofline']['frac'])"
