#!/bin/bash
# two-kernel crystal path: unroll of the branch-free per-system loop (ILP across systems) x blocks/SM
run() { echo "== $1"; python tools/bench_all.py --reps 6 --only crystal 2>&1 | grep stepB | cut -c1-120; }
for cfg in "1 6" "2 6" "2 5" "3 5" "3 4" "4 4"; do
  set -- $cfg
  F="-DMM_CRYSTAL_UNROLL=$1 -DMM_CRYSTAL_MINBLOCKS=$2"
  MM_NVCC_EXTRA="$F" python materialmodels.jl_b200/build.py -v 2>&1 | grep -A2 "crystal_newton_kernelILb0ELb1ELi0" | grep -E "Used|spill" | tr '\n' ' '; echo
  run "$F"
done
python materialmodels.jl_b200/build.py > /dev/null 2>&1
