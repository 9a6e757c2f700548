#!/bin/bash
for u in 1 2 4; do
  MM_NVCC_EXTRA="-DMM_CELLS_KE_UNROLL=$u" python materialmodels.jl_b200/build.py -v 2>&1 | grep -A2 "cells_kernelILi1ELi8ELi8ELi128ELb1" | grep -E "Used|spill" | tr '\n' ' '; echo
  echo "== MM_CELLS_KE_UNROLL=$u"; python tools/bench_all.py --reps 5 --only cells 2>&1 | grep "nq=8 fe+state+ke" | cut -c1-200
done
python materialmodels.jl_b200/build.py > /dev/null 2>&1
