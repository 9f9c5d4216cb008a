for v in 0 4 8; do
EVE_B200_IN_STREAM_CS_BIG=$v timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-extra --no-e2e > gpurun_out/r2bi_bench_$v.json 2> gpurun_out/r2bi_bench_$v.err; echo "bench cs_big=$v rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2bi_bench_$v.json')); print(j['ms_per_step'], j['value'], j['clocks']['sm_mhz'], j['clocks']['power_w_max'])"
done
