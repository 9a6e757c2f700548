#!/bin/bash
for mb in 5 6 8; do
  MM_NVCC_EXTRA="-DMM_WRAPTI_MINBLOCKS=$mb" python materialmodels.jl_b200/build.py > /dev/null 2>&1
  echo "== MM_WRAPTI_MINBLOCKS=$mb"; python tools/bench_all.py --reps 6 --only wrapti 2>&1 | cut -c1-170
done
python materialmodels.jl_b200/build.py > /dev/null 2>&1
