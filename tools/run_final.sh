# The commands behind profiles/r01_*: run on one B200 with `gpurun -- 'bash tools/run_final.sh'`.
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; cut -c1-600 gpurun_out/bench_r1_n1.json; tail -2 gpurun_out/bench_r1_n1.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_r1_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r1_reference.json
# launch list of the bench command and one full capture of the dominant kernel (numbers under ncu are not bench values)
if [ "${WITH_NCU:-0}" = "1" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
  timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_fold -s 5 -c 3 -o gpurun_out/prof_fold_r1_final2 -f python tools/diag_steps.py 9 > gpurun_out/ncu_final2.log 2>&1; tail -2 gpurun_out/ncu_final2.log
fi
