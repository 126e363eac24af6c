BW_SUB_ROWS=1024 timeout 400 python -m pytest tests/test_gpu_multi.py -x -q --timeout 300 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29561 bench.py --gpus 2 --steps 30 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_r1_n2.json 2> gpurun_out/bench_r1_n2.err; cut -c1-300 gpurun_out/bench_r1_n2.json; tail -3 gpurun_out/bench_r1_n2.err
BW_SUB_ROWS=8388608 timeout 200 $TR --master-port 29562 bench.py --gpus 2 --steps 30 --warmup 3 --no-e2e --no-cpu 2>&1 | grep -E "^\{|rror" | cut -c1-900
