timeout 300 python -m pytest tests/test_gpu_models.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/prof_step.py refine > gpurun_out/prof_step2.txt 2>&1
grep -v "^/opt\|_warn" gpurun_out/prof_step2.txt | head -60 | cut -c1-150
