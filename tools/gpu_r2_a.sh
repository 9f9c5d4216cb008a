# round 2, step A: fused InstanceNorm kernels -- op tests, full GPU suite, bench A/B
start=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_in_fused.py -q -x > gpurun_out/r2a_infused.txt 2>&1
echo "in_fused rc=$? secs=$(( $(date +%s) - start ))"; tail -15 gpurun_out/r2a_infused.txt
start=$(date +%s)
timeout 600 python -m pytest tests/ -q -m gpu > gpurun_out/r2a_pytest.txt 2>&1
echo "pytest rc=$? secs=$(( $(date +%s) - start ))"; tail -15 gpurun_out/r2a_pytest.txt
start=$(date +%s)
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2a_bench_fused.json 2> gpurun_out/r2a_bench_fused.err
echo "bench fused rc=$? secs=$(( $(date +%s) - start ))"; cut -c1-400 gpurun_out/r2a_bench_fused.json
EVE_B200_FUSED_NORM=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_legacy.json 2> gpurun_out/r2a_bench_legacy.err
echo "bench legacy rc=$?"; cut -c1-400 gpurun_out/r2a_bench_legacy.json
timeout 200 python tools/prof_step.py refine > gpurun_out/r2a_prof_step.txt 2>&1
echo "prof rc=$?"; head -40 gpurun_out/r2a_prof_step.txt
