timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -5
echo "== default 1 GPU"
BW_TIMING=1 timeout 200 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu 2>&1 | grep -E "avg|^\{|rror" | cut -c1-250
echo "== stride 8, 125k keys"
BW_TIMING=1 timeout 200 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu --n-keys 125000 --ts-stride 8 2>&1 | grep -E "avg|^\{|rror" | cut -c1-250
echo "== stride 2, 500k keys"
BW_TIMING=1 timeout 200 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu --n-keys 500000 --ts-stride 2 2>&1 | grep -E "avg|^\{|rror" | cut -c1-250
