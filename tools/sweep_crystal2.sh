#!/bin/bash
for mb in 4 5 6; do
  MM_NVCC_EXTRA="-DMM_CRYSTAL_MINBLOCKS=$mb" python materialmodels.jl_b200/build.py > /dev/null 2>&1
  echo "== MINBLOCKS=$mb"; python tools/bench_all.py --reps 4 --only crystal 2>&1 | cut -c1-110
done
python materialmodels.jl_b200/build.py > /dev/null 2>&1
