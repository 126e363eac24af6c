timeout 400 python -m pytest tests/test_gpu_multi.py -x -q --timeout 300 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
BW_TIMING=1 timeout 200 $TR --master-port 29561 bench.py --gpus 2 --steps 30 --warmup 3 --no-e2e --no-cpu 2>&1 | grep -E "rank 0\]|^\{|rror" | cut -c1-300
