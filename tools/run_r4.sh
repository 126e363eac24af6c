timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 300 2>&1 | tail -5
echo "== default 1 GPU"
BW_TIMING=1 timeout 200 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu 2>&1 | grep -E "avg|^\{|rror" | cut -c1-250
echo "== BW_SEG=1"
BW_SEG=1 BW_TIMING=1 timeout 200 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu 2>&1 | grep -E "avg|rror" | cut -c1-250
for auto in 0 1; do
  echo "== stride 8, 125k keys, BW_SUB_AUTO=$auto"
  BW_SUB_AUTO=$auto BW_TIMING=1 timeout 200 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu --n-keys 125000 --ts-stride 8 2>&1 | grep -E "avg|^\{|rror" | cut -c1-250
done
echo "== e2e diag"
timeout 120 python tools/diag_e2e.py 16 2>&1 | tail -6
