#!/bin/bash
# TMA-staged AoS kernels: 1 vs 2 input images (prefetch of the next tile before the math)
for cfg in "1 64 128" "2 64 128" "2 32 128" "2 64 64" "2 128 256"; do set -- $cfg
  MM_NVCC_EXTRA="-DMM_AOS_NBUF=$1 -DMM_PLASTIC_AOS_BS=$2 -DMM_STREAM_AOS_BS=$3" python materialmodels.jl_b200/build.py > /dev/null 2>&1
  echo "== NBUF=$1 plastic BS=$2 streaming BS=$3"
  python tools/bench_all.py --reps 5 --only plastic,hyper,xu,transviso 2>&1 | grep -E " aos" | cut -c1-125
done
python materialmodels.jl_b200/build.py > /dev/null 2>&1
