for cfg in "102400 8" "102400 16" "65536 16" "200000 16"; do
  set -- $cfg
  echo "=== smem=$1 maxcs=$2"
  EVE_B200_IN_SMEM=$1 EVE_B200_IN_MAXCS=$2 timeout 200 python tools/bench_in.py all 2>&1 | grep -E "total|9216x16 |9216x64|2304x64 |576x128|1024x64|9216x32|2304x128"
done
