#!/usr/bin/env python3
"""Chunk-size sweep of the host-buffer J2 path (dense and factored-compact returns), pinned buffers, one GPU.
usage: python tools/e2e_sweep.py [points] [full]   -> JSON lines: {"form", "chunk", "points", "updates_per_s", "GBps_h2d", "GBps_d2h"}"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import materialmodels_jl_b200 as mm
import materialmodels_jl_b200._lib as L
from materialmodels_jl_b200 import synth as sy

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2 * 10**7
lib = L.load()
m = mm.Plastic(**sy.PLASTIC_PARAMS)
ncA, loA, hiA, seedA = sy.c2_strain(0)
ncB, loB, hiB, seedB = sy.c2_strain(1)
epsA = mm.synth_uniform(N, ncA, seedA, loA, hiA)
_, _, stA = mm.material_response(m, epsA, mm.initial_material_state(m, N), tangent=False)
epsB = mm.synth_uniform(N, ncB, seedB, loB, hiB)
h_eps = mm.PinnedArray((6, N)).array
h_st = mm.PinnedArray((14, N)).array
for lo in range(0, N, 1 << 24):
    hi = min(N, lo + (1 << 24))
    h_eps[:, lo:hi] = epsB[:, lo:hi].cpu().numpy()
    h_st[:, lo:hi] = stA.data[:, lo:hi].cpu().numpy()
del epsA, epsB, stA
torch.cuda.empty_cache()
outs = {"stress": mm.PinnedArray((6, N)).array, "flags": mm.PinnedArray((N,), np.uint8).array, "iters": mm.PinnedArray((N,), np.uint16).array,
        "rows": mm.PinnedArray((int(0.54 * N) + 1024, 27)).array}
for chunk in (1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22):
    lib.mmh_set_chunk_points(chunk)
    ts = []
    for rep in range(4):
        t0 = time.perf_counter()
        _, T, _ = mm.material_response(m, h_eps, mm.PlasticState(h_st), tangent="factored", out=outs)
        ts.append(time.perf_counter() - t0)
    dt = min(ts[1:])
    print(json.dumps({"form": "factored", "chunk": chunk, "points": N, "updates_per_s": N / dt, "first_call_updates_per_s": N / ts[0],
                      "GBps_h2d": 160 * N / dt / 1e9, "GBps_d2h": (51 * N + 216 * T.n_rows) / dt / 1e9}), flush=True)
