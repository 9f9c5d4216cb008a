timeout 120 python tools/probe_mma.py > gpurun_out/r2ao_probe_mma.txt 2>&1; echo "rc=$?"; tail -14 gpurun_out/r2ao_probe_mma.txt
