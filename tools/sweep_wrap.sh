#!/bin/bash
# wrapped J2 kernels: register cap (min blocks of 128 threads per SM)
for mb in 2 3 4; do
  MM_NVCC_EXTRA="-DMM_WRAP_MINBLOCKS=$mb" python materialmodels.jl_b200/build.py -v 2>&1 | grep -A2 "soa_kernelINS_16WrappedPlasticOpILi4" | grep -E "Used|spill" | tr '\n' ' '; echo
  echo "== MM_WRAP_MINBLOCKS=$mb"; python tools/bench_all.py --reps 4 --only wrapped 2>&1 | cut -c1-250
done
python materialmodels.jl_b200/build.py > /dev/null 2>&1
