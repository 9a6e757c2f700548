#!/bin/bash
# crystal kernel: unroll factor of the branch-free per-system loop x min blocks/SM  (run on the GPU box)
for u in 1 2 3 4 6 12; do
 for mb in 4 5; do
  MM_NVCC_EXTRA="-DMM_CRYSTAL_UNROLL=$u -DMM_CRYSTAL_MINBLOCKS=$mb" python materialmodels.jl_b200/build.py -v 2>&1 | grep -A2 "crystal_persistent_kernelILb0ELi0" | grep -E "Used|spill" | tr '\n' ' '
  echo; echo "== UNROLL=$u MINBLOCKS=$mb"; python tools/bench_all.py --reps 4 --only crystal 2>&1 | cut -c1-120
 done
done
python materialmodels.jl_b200/build.py > /dev/null 2>&1
