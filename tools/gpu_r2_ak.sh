timeout 900 python -m pytest tests/ -q -m gpu > gpurun_out/r2f_pytest.txt 2>&1; tail -4 gpurun_out/r2f_pytest.txt
