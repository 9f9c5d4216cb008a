for cfg in "1 0" "1 32" "2 32" "2 0"; do
set -- $cfg
EVE_B200_TC_CTAS=$1 EVE_B200_TC_KC=$2 timeout 600 python tools/conv_table.py > gpurun_out/r2at_conv_table_$1_$2.txt 2>&1; echo "ctas=$1 kc=$2: $(head -1 gpurun_out/r2at_conv_table_$1_$2.txt)"
done
