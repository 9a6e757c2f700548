#!/bin/bash
# usage: run_variants.sh "<bench_all section>" "<grep pattern>" tag1 tag2 ...   (variants built by tools/build_variants.py)
sec=$1; pat=$2; shift 2
cp materialmodels.jl_b200/libmmb200.so /tmp/libmmb200_orig.so
for t in "$@"; do
  cp variants/libmmb200_$t.so materialmodels.jl_b200/libmmb200.so
  echo "== $t"; python tools/bench_all.py --reps 8 --only $sec 2>&1 | grep -E "$pat" | cut -c1-170
done
cp /tmp/libmmb200_orig.so materialmodels.jl_b200/libmmb200.so
