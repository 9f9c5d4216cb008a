timeout 120 python tools/probe_mma.py > gpurun_out/r2y_probe_mma.txt 2>&1; tail -16 gpurun_out/r2y_probe_mma.txt
