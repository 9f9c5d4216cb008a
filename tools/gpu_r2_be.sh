timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_options.py -q -x -m gpu -k "upsample or halo_row or tc_row_wgrad or stem_fused" > gpurun_out/r2be_pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2be_pytest.txt
for v in 2 1; do
EVE_B200_TC_ROW_WGRAD=$v timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extra --no-e2e > gpurun_out/r2be_bench_$v.json 2> gpurun_out/r2be_bench_$v.err; echo "bench row_wgrad=$v rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2be_bench_$v.json')); print(j['ms_per_step'], j['value'], j['roofline']['frac'], j['roofline']['per_kind'])"
done
