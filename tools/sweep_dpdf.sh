#!/bin/bash
for mb in 2 3 4; do
  MM_NVCC_EXTRA="-DMM_DPDF_MINBLOCKS=$mb" python materialmodels.jl_b200/build.py -v 2>&1 | grep -A2 "soa_kernelINS_7HyperOpILi0ELi1ELi2" | grep -E "Used|spill" | tr '\n' ' '; echo
  echo "== MM_DPDF_MINBLOCKS=$mb"; python tools/bench_all.py --reps 6 --only traits 2>&1 | grep dPdF | cut -c1-160
done
python materialmodels.jl_b200/build.py > /dev/null 2>&1
