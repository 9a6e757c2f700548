#!/bin/bash
# two-kernel crystal path, 66-slot scratch: Newton-kernel block size x min blocks/SM  (run on the GPU box)
run() { echo "== $1 | THRESH=${MM_CRYSTAL_THRESH:-4}"; python tools/bench_all.py --reps 4 --only crystal 2>&1 | cut -c1-120; }
for cfg in "64 5" "64 6" "32 10" "32 12" "32 14"; do
  set -- $cfg
  F="-DMM_CRYSTAL_BS=$1 -DMM_CRYSTAL_MINBLOCKS=$2"
  MM_NVCC_EXTRA="$F" python materialmodels.jl_b200/build.py -v 2>&1 | grep -A2 "crystal_newton_kernelILb0ELb1ELi0" | grep -E "Used|spill" | tr '\n' ' '; echo
  run "$F"
done
F="-DMM_CRYSTAL_BS=32 -DMM_CRYSTAL_MINBLOCKS=12 -DMM_CRYSTAL_FINISH_BS=128 -DMM_CRYSTAL_FINISH_MINBLOCKS=3"
MM_NVCC_EXTRA="$F" python materialmodels.jl_b200/build.py > /dev/null 2>&1; run "$F"
python materialmodels.jl_b200/build.py > /dev/null 2>&1
