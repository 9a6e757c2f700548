#!/bin/bash
# tuning sweep on the GPU box for the lane-persistent crystal kernel
echo "== non-persistent (one point per thread, rolled loops)"; MM_CRYSTAL_PERSISTENT=0 python tools/bench_all.py --reps 4 --only crystal 2>&1 | cut -c1-120
for th in 4 8 16 24 32; do echo "== persistent THRESH=$th"; MM_CRYSTAL_THRESH=$th python tools/bench_all.py --reps 4 --only crystal 2>&1 | cut -c1-120; done
for ch in 128 4096; do echo "== persistent THRESH=16 CHUNK=$ch"; MM_CRYSTAL_THRESH=16 MM_CRYSTAL_CHUNK=$ch python tools/bench_all.py --reps 4 --only crystal 2>&1 | cut -c1-120; done
