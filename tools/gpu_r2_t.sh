for mb in 8 24 48; do
  echo "=== chunk $mb MB"
  BENCH_IN_CHUNK_MB=$mb BENCH_IN_SHAPES=240x9216x64,240x9216x16,240x2304x64,480x1024x64,240x576x128 timeout 200 python tools/bench_in.py all 2>&1 | tail -12
done
