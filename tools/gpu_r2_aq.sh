for m in 0 1; do
EVE_B200_TC_PAIR=$m timeout 600 python tools/conv_table.py > gpurun_out/r2aq_conv_table_$m.txt 2>&1; head -1 gpurun_out/r2aq_conv_table_$m.txt
done
