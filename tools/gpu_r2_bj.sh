timeout 400 python -m pytest tests/test_gpu_options.py tests/test_gpu_bench_size.py -q -x -m gpu -k "halo_row or tc_row_wgrad or conv_passes_at_bench" > gpurun_out/r2bj_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2bj_pytest.txt
timeout 200 python tools/conv_table.py > gpurun_out/r2bj_conv_table.txt 2>&1; head -1 gpurun_out/r2bj_conv_table.txt; grep "^wgrad  *240  *72  *128" gpurun_out/r2bj_conv_table.txt
timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-extra --no-e2e > gpurun_out/r2bj_bench.json 2> gpurun_out/r2bj_bench.err; echo "bench rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2bj_bench.json')); print(j['ms_per_step'], j['value'], j['roofline']['frac'], j['roofline']['per_kind']['conv_wgrad'])"
