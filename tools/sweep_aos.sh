#!/bin/bash
for cfg in "64 128" "64 64" "128 256" "32 128"; do set -- $cfg
  MM_NVCC_EXTRA="-DMM_PLASTIC_AOS_BS=$1 -DMM_STREAM_AOS_BS=$2" python materialmodels.jl_b200/build.py > /dev/null 2>&1
  echo "== plastic AoS BS=$1, streaming AoS BS=$2"
  python tools/bench_all.py --reps 4 --only plastic,hyper,xu 2>&1 | grep -E " aos" | cut -c1-110
done
python materialmodels.jl_b200/build.py > /dev/null 2>&1
