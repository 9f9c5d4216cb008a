timeout 120 python tools/debug_dual.py > gpurun_out/r2az_debug.txt 2>&1; echo "rc=$?"; grep "diff\|timeout\|Error" gpurun_out/r2az_debug.txt | cut -c1-200
EVE_B200_TC_DUAL=1 timeout 150 python tools/conv_table.py > gpurun_out/r2az_ct.txt 2>&1; echo "rc=$?"; head -1 gpurun_out/r2az_ct.txt | cut -c1-200
