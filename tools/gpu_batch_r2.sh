set -x
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum
ncu --csv --clock-control none --metrics $M -k regex:crystal --launch-skip 3 --launch-count 3 python tools/profile_one.py crystal 2e6 > gpurun_out/r2_crystal_fp64_metrics.csv 2> gpurun_out/ncu_d.err
ncu --set full --clock-control none --import-source on -k regex:crystal_newton --launch-skip 1 --launch-count 1 -o gpurun_out/r2_crystal_final python tools/profile_one.py crystal 2e6 > gpurun_out/ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:soa_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/r2_plastic_soa python tools/profile_one.py plastic 1e7 > gpurun_out/ncu_f.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
tail -3 gpurun_out/r2_bench_c.err
