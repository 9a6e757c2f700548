# Xu SoA residency-cap variants, alternated, each timed right after the hyperelastic kernels (the bench's order), 15 reps (sustained state)
cp materialmodels.jl_b200/libmmb200.so /tmp/libmmb200_orig.so
for round in 1 2; do
for t in base cap2 cap3 cap4 cap6; do
  cp variants/libmmb200_$t.so materialmodels.jl_b200/libmmb200.so
  echo "== $t (round $round)"; python tools/bench_all.py --reps 15 --only hyper,xu 2>&1 | grep -E "xu_needleman 3D soa" | cut -c1-150
done
done
cp /tmp/libmmb200_orig.so materialmodels.jl_b200/libmmb200.so
