timeout 900 python -m pytest tests/ -q -m gpu > gpurun_out/r2f_pytest.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2f_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
