#!/usr/bin/env python3
"""cProfile of the Python mirror's per-call overhead (N = 1000 J2 points, 3000 eager calls)."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import materialmodels_jl_b200 as mm
from materialmodels_jl_b200 import synth as sy

m = mm.Plastic(**sy.PLASTIC_PARAMS)
N = 1000
e = mm.synth_uniform(N, 6, 1, -7e-4, 7e-4)
st = mm.initial_material_state(m, N)
opt = {"defer_check": True}
for _ in range(100):
    mm.material_response(m, e, st, options=opt)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3000):
    mm.material_response(m, e, st, options=opt)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
