# Per-kernel evidence (run under gpurun on one B200; every command under its own timeout).  Reports land in gpurun_out/;
# tools/ncu_summary.py turns them into the markdown kept under profiles/.
set +e
N="ncu --set full --clock-control none --import-source on -f"
# 1. launch list of the bench command itself (the kernel's SHARE of the step must agree with the live CUDA-event numbers)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu --min-timed-s 0 > /dev/null 2>&1
# 2. the streaming path: scatter, verdict, segment fold, spill, K4 (steady-state activations: skip the first launches)
timeout 300 $N -k regex:'k_scatter|k_segfold|k_verdict|k_close_dirty|k_spill' -s 10 -c 10 -o gpurun_out/prof_stream_r2 python tools/diag_steps.py 8 > gpurun_out/ncu_stream.log 2>&1; tail -1 gpurun_out/ncu_stream.log
# 3. activations with late rows: the suspect table kernels
timeout 300 $N -k regex:'k_late_' -s 7 -c 7 -o gpurun_out/prof_late_r2 python bench.py --steps 4 --warmup 3 --late-frac 0.01 --no-e2e --no-cpu --min-timed-s 0 > gpurun_out/ncu_late.log 2>&1; tail -1 gpurun_out/ncu_late.log
# 4. the direct path (BW_STREAM=0): lateness pass + hash-table fold
BW_STREAM=0 timeout 300 $N -k regex:'k_fold|k_prepass' -s 9 -c 6 -o gpurun_out/prof_direct_r2 python tools/diag_steps.py 6 > gpurun_out/ncu_direct.log 2>&1; tail -1 gpurun_out/ncu_direct.log
# 5. the row exchange + sort path, two loopback ranks on this one GPU
timeout 300 $N -k regex:'k_part_|k_slow_' -s 6 -c 10 -o gpurun_out/prof_xchg_r2 python tools/diag_loopback.py > gpurun_out/ncu_xchg.log 2>&1; tail -1 gpurun_out/ncu_xchg.log
# 6. K5 / K6
timeout 200 $N -k regex:'k_smap_eval|k_smap_update|k_smap_slots|k_keyed_heads' -s 4 -c 4 -o gpurun_out/prof_smap_r2 python bench.py --config c2 --steps 2 --warmup 1 --batch-rows 1048576 > gpurun_out/ncu_smap.log 2>&1; tail -1 gpurun_out/ncu_smap.log
timeout 200 $N -k regex:'k_join_apply|k_keyed_heads' -s 2 -c 3 -o gpurun_out/prof_join_r2 python bench.py --config c4 --steps 2 --warmup 1 --batch-rows 524288 > gpurun_out/ncu_join.log 2>&1; tail -1 gpurun_out/ncu_join.log
# what travels back (gpurun_out is capped at 64 MiB): the raw metric pages as CSV for every report, the full report of the
# streaming kernels only (source view of the two kernels that matter)
for r in stream late direct xchg smap join; do
  ncu -i gpurun_out/prof_${r}_r2.ncu-rep --page raw --csv > gpurun_out/raw_${r}_r2.csv 2>/dev/null
  [ $r != stream ] && rm -f gpurun_out/prof_${r}_r2.ncu-rep
done
ls -la gpurun_out/
