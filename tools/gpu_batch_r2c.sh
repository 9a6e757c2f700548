# final batch of round 2: full GPU suite, bench at HEAD, ncu launch list of the same bench command (reduced steps), crystal captures
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_gputests_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches_final.csv python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/bench_under_ncu.log 2>&1
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum
ncu --csv --metrics $M --clock-control none -k regex:crystal --launch-skip 4 --launch-count 4 python tools/profile_one.py crystal 2e6 > gpurun_out/r2_crystal_fp64_metrics_aset.csv 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:crystal_newton -s 6 -c 2 -o gpurun_out/crystal_aset_final python tools/profile_one.py crystal 2e6 > /dev/null 2>&1
tail -c 600 gpurun_out/bench_final.json
