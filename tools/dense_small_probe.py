#!/usr/bin/env python3
"""Host-buffer J2 call with a dense tangent at small and medium batch sizes: factored rows + host expansion vs plain dense transfer."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import materialmodels_jl_b200 as mm
import materialmodels_jl_b200._lib as L
from materialmodels_jl_b200 import synth as sy
lib = L.load()
m = mm.Plastic(**sy.PLASTIC_PARAMS)
for N in (1000, 10000, 100000, 300000, 1000000, 4000000):
    nc, lo, hi, seed = sy.c2_strain(1)
    eps = mm.PinnedArray((6, N)).array; eps[:] = sy.host_uniform(N, nc, seed, lo, hi)
    st = mm.PlasticState(mm.PinnedArray((14, N)).array); st.data[:] = 0.0
    out = {"stress": mm.PinnedArray((6, N)).array, "tangent": mm.PinnedArray((36, N)).array, "state": mm.PinnedArray((14, N)).array,
           "flags": mm.PinnedArray((N,), np.uint8).array, "iters": mm.PinnedArray((N,), np.uint16).array}
    res = {}
    for via in (1, 0):
        prev = lib.mm_set_tuning(b"mmh_plastic_via_factored", via)
        ts = []
        for r in range(12):
            t0 = time.perf_counter(); mm.material_response(m, eps, st, out=out, options={"defer_check": True}); ts.append(time.perf_counter() - t0)
        lib.mm_set_tuning(b"mmh_plastic_via_factored", prev)
        res["via_factored" if via else "plain"] = round(1e6 * sorted(ts[2:])[len(ts[2:]) // 2], 1)
    print(json.dumps({"N": N, "us_per_call": res}), flush=True)
