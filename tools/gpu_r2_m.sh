timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_pytest.txt 2>&1; tail -6 gpurun_out/r2m_pytest.txt
timeout 1500 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2m_bench.json; tail -2 gpurun_out/r2m_bench.err
