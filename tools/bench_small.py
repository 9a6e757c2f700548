#!/usr/bin/env python3
"""Small-batch regime of the J2 update: time per step for N = 1e3 .. 1e7 points, (a) eager calls through the Python mirror
(ctypes + torch allocations per call), (b) the same 20 chained steps captured once in a CUDA graph and replayed.
One JSON line per N.  No oracle use."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import materialmodels_jl_b200 as mm
from materialmodels_jl_b200 import synth as sy

DEFER = {"defer_check": True}
STEPS = 20
m = mm.Plastic(**sy.PLASTIC_PARAMS)
for N in (1000, 10000, 100000, 1000000, 10000000):
    eps = [mm.synth_uniform(N, 6, 700 + k, -6.6e-4 * (1 + 0.05 * k), 6.6e-4 * (1 + 0.05 * k)) for k in range(STEPS)]

    def loop():
        st = mm.initial_material_state(m, N)
        for e in eps:
            s, t, st = mm.material_response(m, e, st, options=DEFER)
        return s, t, st

    loop(); torch.cuda.synchronize()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        loop()
    torch.cuda.synchronize()
    eager_us = (time.perf_counter() - t0) / (reps * STEPS) * 1e6
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = loop()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record(); torch.cuda.synchronize()
    graph_us = a.elapsed_time(b) / (reps * STEPS) * 1e3
    print(json.dumps({"N": N, "steps_per_graph": STEPS, "eager_us_per_step": round(eager_us, 2), "graph_us_per_step": round(graph_us, 2),
                      "eager_updates_per_s": N / eager_us * 1e6, "graph_updates_per_s": N / graph_us * 1e6,
                      "graph_frac_of_hbm_peak": round(608.0 * N / (graph_us * 1e-6) / 1e9 / 6454.0, 3)}), flush=True)
    del eps, out, g
    torch.cuda.empty_cache()
