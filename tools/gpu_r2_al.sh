for cs in 0 4 2 1; do
EVE_B200_IN_STREAM_CS=$cs BENCH_IN_STREAM=2 timeout 300 python tools/bench_in.py bwd > gpurun_out/r2al_in_bwd_cs$cs.txt 2>&1
done
paste <(cut -c1-42 gpurun_out/r2al_in_bwd_cs0.txt) <(cut -c19-42 gpurun_out/r2al_in_bwd_cs4.txt) <(cut -c19-42 gpurun_out/r2al_in_bwd_cs2.txt) <(cut -c19-42 gpurun_out/r2al_in_bwd_cs1.txt)
for cs in 4 2; do
EVE_B200_IN_STREAM_CS=$cs timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2al_bench_cs$cs.json 2> gpurun_out/r2al_bench_cs$cs.err; cut -c1-230 gpurun_out/r2al_bench_cs$cs.json
done
