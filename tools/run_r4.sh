echo "== default 1 GPU"
BW_TIMING=1 timeout 200 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu 2>&1 | grep -E "avg|^\{|rror" | cut -c1-250
for auto in 0 1; do
  echo "== stride 8, 125k keys, BW_SUB_AUTO=$auto"
  BW_SUB_AUTO=$auto BW_TIMING=1 timeout 200 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu --n-keys 125000 --ts-stride 8 2>&1 | grep -E "avg|^\{|rror" | cut -c1-250
done
echo "== stride 4, 250k keys, auto"
BW_TIMING=1 timeout 200 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu --n-keys 250000 --ts-stride 4 2>&1 | grep -E "avg|^\{|rror" | cut -c1-250
