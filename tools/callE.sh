start=$(date +%s)
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_full.txt 2>&1
echo "pytest rc=$? secs=$(( $(date +%s) - start ))"
tail -6 gpurun_out/pytest_full.txt
start=$(date +%s)
timeout 600 python bench.py > gpurun_out/bench_rowk.json 2> gpurun_out/bench_rowk.err
echo "bench rc=$? secs=$(( $(date +%s) - start ))"
cat gpurun_out/bench_rowk.json | cut -c1-1500
