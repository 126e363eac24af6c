BW_SEG=1 BW_TIMING=1 timeout 200 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu 2>&1 | grep -E "avg|rror" | cut -c1-200
BW_SEG=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fold_seg --launch-skip 12 --launch-count 1 -o gpurun_out/prof_seg -f python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_seg.log 2>&1
tail -3 gpurun_out/ncu_seg.log
