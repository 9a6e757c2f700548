#!/bin/bash
for mb in 2 3 4; do
  MM_NVCC_EXTRA="-DMM_CELLS_MINBLOCKS=$mb" python materialmodels.jl_b200/build.py -v 2>&1 | grep -A2 "cells_kernelILi1ELi8ELi8ELi128ELb0" | grep -E "Used|spill" | tr '\n' ' '; echo
  echo "== MM_CELLS_MINBLOCKS=$mb"; python tools/bench_all.py --reps 5 --only cells 2>&1 | grep -v "state+ke" | cut -c1-200
done
python materialmodels.jl_b200/build.py > /dev/null 2>&1
