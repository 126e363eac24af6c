TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
BW_TIMING=1 BW_NO_OVERLAP=1 timeout 150 $TR --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/n8_serial.log 2>&1
BW_FOLD_FULL=1 timeout 150 $TR --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/n8_full.log 2>&1
timeout 150 $TR --master-port 29553 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --no-cpu --exchange nccl > gpurun_out/n8_nccl.log 2>&1
grep -h "avg\|^{" gpurun_out/n8_serial.log | cut -c1-260 | head -60
grep -h "^{" gpurun_out/n8_full.log gpurun_out/n8_nccl.log | cut -c1-260
