for pl in 0 1; do
EVE_B200_IN_PLAN=$pl timeout 300 python tools/bench_in.py fwd > gpurun_out/r2an_in_fwd_p$pl.txt 2>&1
EVE_B200_IN_PLAN=$pl BENCH_IN_STREAM=0 timeout 300 python tools/bench_in.py bwd > gpurun_out/r2an_in_bwd_staged_p$pl.txt 2>&1
done
paste <(cut -c1-42 gpurun_out/r2an_in_fwd_p0.txt) <(cut -c19-42 gpurun_out/r2an_in_fwd_p1.txt)
paste <(cut -c1-42 gpurun_out/r2an_in_bwd_staged_p0.txt) <(cut -c19-42 gpurun_out/r2an_in_bwd_staged_p1.txt)
for pl in 0 1; do
EVE_B200_IN_PLAN=$pl timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2an_bench_p$pl.json 2> gpurun_out/r2an_bench_p$pl.err; cut -c1-230 gpurun_out/r2an_bench_p$pl.json
done
