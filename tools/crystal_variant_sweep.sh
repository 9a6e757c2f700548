cp materialmodels.jl_b200/libmmb200.so /tmp/libmmb200_orig.so
for t in base u2 u3 mb5 bs32 bs128; do
  cp variants/libmmb200_$t.so materialmodels.jl_b200/libmmb200.so
  echo "== $t"; python tools/crystal_sweep.py 1e7 0,0 2>&1 | tail -1
done
cp variants/libmmb200_base.so materialmodels.jl_b200/libmmb200.so
echo "== knobs (base)"; python tools/crystal_sweep.py 1e7 1,0 2,0 3,0 4,0 6,0 8,0 12,0 4,128 4,256 4,512 4,2048 2,256 2>&1 | tail -12
cp /tmp/libmmb200_orig.so materialmodels.jl_b200/libmmb200.so
