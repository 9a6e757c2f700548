#!/bin/bash
# crystal kernel: small loops rolled/unrolled x min blocks/SM x finalize threshold  (run on the GPU box)
run() { echo "== $1 | THRESH=${MM_CRYSTAL_THRESH:-8}"; python tools/bench_all.py --reps 4 --only crystal 2>&1 | cut -c1-120; }
for us in 1 12; do
 for mb in 3 4; do
  F="-DMM_CRYSTAL_UNROLL_SMALL=$us -DMM_CRYSTAL_MINBLOCKS=$mb"
  MM_NVCC_EXTRA="$F" python materialmodels.jl_b200/build.py -v 2>&1 | grep -A2 "crystal_persistent_kernelILb0ELi0" | grep -E "Used|spill" | tr '\n' ' '; echo
  run "$F"
 done
done
F="-DMM_CRYSTAL_UNROLL_SMALL=1 -DMM_CRYSTAL_MINBLOCKS=4 -DMM_CRYSTAL_BS=128"
MM_NVCC_EXTRA="$F" python materialmodels.jl_b200/build.py > /dev/null 2>&1; run "$F"
F="-DMM_CRYSTAL_UNROLL_SMALL=1 -DMM_CRYSTAL_MINBLOCKS=8 -DMM_CRYSTAL_BS=32"
MM_NVCC_EXTRA="$F" python materialmodels.jl_b200/build.py > /dev/null 2>&1; run "$F"
python materialmodels.jl_b200/build.py > /dev/null 2>&1
for t in 4 12 16; do MM_CRYSTAL_THRESH=$t run "default build"; done
