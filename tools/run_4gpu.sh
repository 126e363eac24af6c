TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29571 bench.py --gpus 4 --steps 30 --warmup 3 --no-e2e --no-cpu 2>gpurun_out/bench_r1_n4.err | grep "^{" > gpurun_out/bench_r1_n4.json; cut -c1-260 gpurun_out/bench_r1_n4.json; grep -i "error\|Traceback" gpurun_out/bench_r1_n4.err | head -5
