export PROBE_STRIP_GEOMS=480x32x32x64x64,240x36x64x64x64,480x16x16x128x128,240x18x32x128x128
for cfg in "0 0" "7 64" "7 32" "5 64" "3 64" "11 64" "15 64"; do
set -- $cfg
echo "== R=$1 KC=$2"
EVE_B200_STRIP_R=$1 EVE_B200_STRIP_KC=$2 timeout 200 python tools/probe_strip.py 5 2>&1 | tail -4 | cut -c1-210
done
