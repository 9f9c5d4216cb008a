timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2s_pytest.txt 2>&1; tail -4 gpurun_out/r2s_pytest.txt
