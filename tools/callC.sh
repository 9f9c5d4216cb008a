timeout 500 python tools/probe_row.py > gpurun_out/probe_row3.txt 2>&1
echo "probe rc=$?"
grep -v "^  .*strips=3\|strips=72" gpurun_out/probe_row3.txt | tail -60
timeout 400 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu > gpurun_out/pytest_phase2.txt 2>&1
echo "pytest rc=$?"
tail -5 gpurun_out/pytest_phase2.txt
