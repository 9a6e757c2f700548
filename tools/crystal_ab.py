#!/usr/bin/env python3
"""A/B timing of the crystal-plasticity Newton kernel variants (mm_set_tuning "crystal_variant": 0 = dense branch-free pass,
1 = sparse per-lane active-list pass, 2 = phased pass [default]) on BASELINE.json configs[3] (FCC, 1e7 points, step B), and how many points differ bit-wise between the dense and the phased variant.
usage: python tools/crystal_ab.py [points] [layout]"""
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
layout = sys.argv[2] if len(sys.argv) > 2 else "soa"
lay = 0 if layout == "soa" else 1
lib = L.load()
cm = mm.CrystalViscoPlastic(slipsystems=mm.slipsystems(mm.FCC(), sy.CRYSTAL_RODRIGUES), **sy.CRYSTAL_PARAMS)
opts = {"defer_check": True}


def run(variant):
    lib.mm_set_tuning(b"crystal_variant", variant)
    st = mm.initial_material_state(cm, N, layout=layout)
    res = {}
    for step in (0, 1):
        nc, lo, hi, seed = sy.c4_strain_increment(step)
        de = mm.synth_uniform(N, nc, seed, lo, hi, layout=layout)
        for _ in range(2):
            out = mm.material_response(cm, de, st, 1.0, options=opts)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = mm.material_response(cm, de, st, 1.0, options=opts)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        res[step] = (ts[len(ts) // 2], out)
        st = out[2]
    return res


r0 = run(0)
extra = {v: run(v) for v in (3, 4, 5, 6)}       # chunks re-ordered by first-pass candidate count: sparse / dense pass; active-set pass with / without re-ordering
rs, r1 = run(1), run(2)
for step in (0, 1):
    o3 = extra[3][step][1]
    print(json.dumps({"step": "AB"[step], "points": N, "layout": layout, "sparse_presorted_ms": round(extra[3][step][0], 3), "dense_presorted_ms": round(extra[4][step][0], 3),
                      "aset_presorted_ms": round(extra[5][step][0], 3), "aset_ms": round(extra[6][step][0], 3),
                      "aset_vs_dense_mask_or_count_differences": int((extra[5][step][1][2].flags != r0[step][1][2].flags).sum().item()) + int((extra[5][step][1][2].iters != r0[step][1][2].iters).sum().item()),
                      "presorted_vs_sparse_bitwise_differences": int((o3[0] != rs[step][1][0]).sum().item()) + int((o3[1] != rs[step][1][1]).sum().item())
                      + int((o3[2].data != rs[step][1][2].data).sum().item()) + int((o3[2].iters != rs[step][1][2].iters).sum().item())}), flush=True)
    (t0, o0), (t1, o1) = r0[step], r1[step]
    tsp = rs[step][0]
    ax = 0 if lay == 0 else 1
    diff = {"sig": int((o0[0] != o1[0]).any(ax).sum().item()), "tan": int((o0[1] != o1[1]).any(ax).sum().item()), "state": int((o0[2].data != o1[2].data).any(ax).sum().item()),
            "mask": int((o0[2].flags != o1[2].flags).sum().item()), "iters": int((o0[2].iters != o1[2].iters).sum().item())}
    rel = float(((o0[1] - o1[1]).abs().max() / o0[1].abs().max()).item())
    print(json.dumps({"step": "AB"[step], "points": N, "layout": layout, "dense_ms": round(t0, 3), "sparse_ms": round(tsp, 3), "phased_ms": round(t1, 3), "speedup_phased_vs_dense": round(t0 / t1, 3),
                      "updates_per_s_phased": N / (t1 * 1e-3), "points_differing_bitwise": diff, "max_rel_tangent_diff": rel,
                      "mean_iters": float(o1[2].iters.float().mean().item()), "n_fail": int(o1[2].n_fail.item())}), flush=True)
lib.mm_set_tuning(b"crystal_variant", -1)
