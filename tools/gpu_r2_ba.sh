for m in 1 0; do
EVE_B200_TC_DUAL=$m timeout 200 python tools/conv_table.py > gpurun_out/r2ba_conv_table_$m.txt 2>&1; head -1 gpurun_out/r2ba_conv_table_$m.txt
EVE_B200_TC_DUAL=$m timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2ba_bench_$m.json 2> gpurun_out/r2ba_bench_$m.err; echo "rc=$?"; cut -c1-230 gpurun_out/r2ba_bench_$m.json
done
