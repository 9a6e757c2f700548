#!/bin/bash
# TMA-staged AoS J2 kernel: tile (= block) size
for bs in 32 64 96 128; do
  MM_NVCC_EXTRA="-DMM_PLASTIC_AOS_BS=$bs" python materialmodels.jl_b200/build.py > /dev/null 2>&1
  echo "== MM_PLASTIC_AOS_BS=$bs"; python tools/bench_all.py --reps 5 --only plastic 2>&1 | grep "stepB aos" | cut -c1-160
done
python materialmodels.jl_b200/build.py > /dev/null 2>&1
