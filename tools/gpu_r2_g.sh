timeout 120 python tools/probe_mma.py > gpurun_out/r2g_probe_mma.txt 2>&1; cat gpurun_out/r2g_probe_mma.txt
