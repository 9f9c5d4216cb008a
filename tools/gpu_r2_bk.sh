for v in 2 4 3; do
EVE_B200_TC_WGRAD_WAVES=$v timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-extra --no-e2e > gpurun_out/r2bk_bench_$v.json 2> gpurun_out/r2bk_bench_$v.err; echo "bench waves=$v rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2bk_bench_$v.json')); print(j['ms_per_step'], j['value'], j['roofline']['per_kind']['conv_wgrad']['ms_per_step'])"
done
