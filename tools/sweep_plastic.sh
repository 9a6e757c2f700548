#!/bin/bash
# tuning sweep on the GPU box: rebuild with different register caps and time the J2 kernel
for mb in 1 3 4 5; do
  MM_NVCC_EXTRA="-DMM_PLASTIC_MINBLOCKS=$mb" python materialmodels.jl_b200/build.py > /dev/null 2>&1
  echo "== MINBLOCKS=$mb"
  python tools/bench_all.py --reps 5 --only plastic 2>&1 | grep -E "stepB soa|all-plastic"
done
python materialmodels.jl_b200/build.py > /dev/null 2>&1
