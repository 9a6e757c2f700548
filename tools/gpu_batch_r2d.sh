# end of round 2: J2 SoA kernel with / without the L2 prefetch hint under ncu (--set full, 1e7 points), launch list of the bench, bench at HEAD
set -x
ncu --set full --clock-control none --import-source on -k regex:soa_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/r2_plastic_soa_prefetch python tools/profile_one.py plastic 1e7 > gpurun_out/ncu_pf.log 2>&1
MM_PLASTIC_SOA_PREFETCH=0 ncu --set full --clock-control none --import-source on -k regex:soa_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/r2_plastic_soa_noprefetch python tools/profile_one.py plastic 1e7 > gpurun_out/ncu_nopf.log 2>&1
python bench.py > gpurun_out/bench_r2_end.json 2> gpurun_out/bench_r2_end.err
tail -3 gpurun_out/bench_r2_end.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches_end.csv python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/bench_under_ncu_end.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_end.json 2> gpurun_out/bench_reference_end.err
head -c 600 gpurun_out/bench_r2_end.json
