timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -4
BW_TIMING=1 timeout 200 python bench.py --steps 60 --warmup 3 --no-e2e --no-cpu 2>&1 | grep -E "avg|^\{|rror" | cut -c1-250
