# Per-kernel evidence for the streaming path (run under gpurun on one B200): phase timers, launch list, full captures.
BW_TIMING=1 timeout 120 python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu 2>&1 >/dev/null | grep bwgpu | tail -4
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r2.csv python tools/diag_steps.py 8 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_scatter|k_segfold' -s 8 -c 2 -o gpurun_out/prof_stream_r2 -f python tools/diag_steps.py 8 > gpurun_out/ncu_stream.log 2>&1; tail -1 gpurun_out/ncu_stream.log
