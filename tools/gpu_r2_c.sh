start=$(date +%s)
timeout 900 python -m pytest tests/ -q -m gpu -x > gpurun_out/r2c_pytest.txt 2>&1
echo "pytest rc=$? secs=$(( $(date +%s) - start ))"; tail -12 gpurun_out/r2c_pytest.txt; grep -E "config[23] (forward|gradient)" gpurun_out/r2c_pytest.txt
export BENCH_IN_SHAPES=240x9216x64,240x2304x64,480x1024x64
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'in_fwd_fused|in_bwd_fused|in_bwd_reduce|in_bwd_apply' -c 12 -o gpurun_out/r2c_in python tools/bench_in.py all > gpurun_out/r2c_ncu.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/r2c_ncu.log
