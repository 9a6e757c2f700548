#!/bin/bash
# two-kernel crystal path: min blocks/SM of the Newton kernel x refill threshold  (run on the GPU box)
run() { echo "== $1 | THRESH=${MM_CRYSTAL_THRESH:-8}"; python tools/bench_all.py --reps 4 --only crystal 2>&1 | cut -c1-120; }
for mb in 4 5; do
  F="-DMM_CRYSTAL_MINBLOCKS=$mb"
  MM_NVCC_EXTRA="$F" python materialmodels.jl_b200/build.py -v 2>&1 | grep -A2 "crystal_newton_kernelILb0ELb1ELi0" | grep -E "Used|spill" | tr '\n' ' '; echo
  for t in 1 2 4 8 12; do MM_CRYSTAL_THRESH=$t run "$F"; done
done
python materialmodels.jl_b200/build.py > /dev/null 2>&1
