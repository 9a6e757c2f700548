#!/usr/bin/env python3
"""Run-time knobs of the crystal Newton kernel (refill threshold, points per warp chunk) on configs[3] step B.
usage: crystal_sweep.py [points]   -> one JSON line per setting"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import materialmodels_jl_b200 as mm
import materialmodels_jl_b200._lib as L
from materialmodels_jl_b200 import synth as sy

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**7
settings = [tuple(map(int, a.split(","))) for a in sys.argv[2:]] or [(0, 0)]
lib = L.load()
cm = mm.CrystalViscoPlastic(slipsystems=mm.slipsystems(mm.FCC(), sy.CRYSTAL_RODRIGUES), **sy.CRYSTAL_PARAMS)
opts = {"defer_check": True}
nc, lo, hi, seed = sy.c4_strain_increment(0)
dA = mm.synth_uniform(N, nc, seed, lo, hi)
_, _, stA = mm.material_response(cm, dA, mm.initial_material_state(cm, N), 1.0, options=opts)
nc, lo, hi, seed = sy.c4_strain_increment(1)
dB = mm.synth_uniform(N, nc, seed, lo, hi)
for thresh, chunk in settings:
    lib.mm_set_tuning(b"crystal_thresh", thresh)
    lib.mm_set_tuning(b"crystal_chunk", chunk)
    for _ in range(2):
        mm.material_response(cm, dB, stA, 1.0, options=opts)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        mm.material_response(cm, dB, stA, 1.0, options=opts)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(json.dumps({"thresh": thresh, "chunk": chunk, "ms": round(ts[2], 3), "ms_min": round(ts[0], 3)}), flush=True)
