#!/bin/bash
# J2 SoA kernel: block size at the same register cap (168 = 3 x 128 threads per SM)
for bs in 64 128 256 384; do
  MM_NVCC_EXTRA="-DMM_PLASTIC_SOA_BS=$bs" python materialmodels.jl_b200/build.py -v 2>&1 | grep -A2 "soa_kernelINS_9PlasticOpELi$bs" | grep -E "Used|spill" | tr '\n' ' '; echo
  echo "== MM_PLASTIC_SOA_BS=$bs"; python tools/bench_all.py --reps 10 --only plastic 2>&1 | grep "stepB soa" | cut -c1-160
done
python materialmodels.jl_b200/build.py > /dev/null 2>&1
