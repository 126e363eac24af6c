# parity of both fold paths, then C1 timings of each (diagnostic)
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 2>&1 | tail -15
for nk in 1000000 125000; do
  for seg in 0 1; do
    echo "== n_keys=$nk BW_SEG=$seg"
    BW_SEG=$seg BW_TIMING=1 timeout 200 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu --n-keys $nk 2>&1 | grep -E "avg|^\{|rror" | cut -c1-330
  done
done
