timeout 400 python tools/probe_row.py > gpurun_out/probe_row.txt 2>&1
echo "probe rc=$?"
EVE_B200_TC_ROW_KERNEL=0 timeout 300 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_ops.py -q -m gpu > gpurun_out/pytest_phase1.txt 2>&1
echo "pytest rc=$?"
tail -5 gpurun_out/pytest_phase1.txt
