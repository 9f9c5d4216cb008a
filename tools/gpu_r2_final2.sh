# final pass of round 2 (second session): tests, whole-step ncu launch list (+ DRAM bytes), per-layer
# table, graph timeline, then the bench line itself
start=$(date +%s)
timeout 900 python -m pytest tests/ -q -m gpu > gpurun_out/r2g_pytest.txt 2>&1
echo "pytest rc=$? secs=$(( $(date +%s) - start ))"; tail -3 gpurun_out/r2g_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
start=$(date +%s)
EVE_BENCH_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 1 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/r2g_launches_bench.log 2>&1
echo "ncu list rc=$? secs=$(( $(date +%s) - start )) lines=$(wc -l < gpurun_out/r2g_launches.csv)"
python tools/ncu_launch_summary.py gpurun_out/r2g_launches.csv > gpurun_out/r2g_launches_summary.txt 2>&1
python tools/ncu_conv_traffic.py gpurun_out/r2g_launches.csv 1 gpurun_out/r2g_conv_traffic.json > /dev/null 2>&1
cp gpurun_out/r2g_conv_traffic.json profiles/conv_traffic.json
gzip -9 gpurun_out/r2g_launches.csv
timeout 300 python tools/conv_table.py > gpurun_out/r2g_conv_table.txt 2>&1; head -1 gpurun_out/r2g_conv_table.txt
timeout 300 python tools/prof_graph.py refine > gpurun_out/r2g_graph_step.txt 2>&1; grep "kernels " gpurun_out/r2g_graph_step.txt
start=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$? secs=$(( $(date +%s) - start ))"; cut -c1-300 gpurun_out/r2g_bench.json
du -sh gpurun_out
