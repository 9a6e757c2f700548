#!/usr/bin/env python3
"""A/B of the L2 prefetch hint of the J2 SoA kernel on BASELINE.json configs[1] (step B): mm_set_tuning "plastic_soa_prefetch" = how many
points ahead lanes 0 and 16 of every warp ask L2 for the 20 input lines (0 = no hint), alternating rounds on one box.  Per distance: the
median of the isolated launches, the mean of a back-to-back run as long as the bench's timed loop (the power-capped figure), and the
number of output words that differ from the run without the hint (must be 0: a hint changes no value).
usage: python tools/plastic_prefetch_ab.py [points] [rounds] [sustained_launches] [distances,comma-separated]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import materialmodels_jl_b200 as mm
import materialmodels_jl_b200._lib as L
from materialmodels_jl_b200 import synth as sy

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**8
ROUNDS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
SUST = int(sys.argv[3]) if len(sys.argv) > 3 else 50
TANGENT = os.environ.get("PF_NO_TANGENT", "0") != "1"      # PF_NO_TANGENT=1: the σ + state launch (no 36-entry tangent written)
VARIANTS = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 14208, 28416, 56832, 113664, 227328]
lib = L.load()
dev = torch.device("cuda", 0)
m = mm.Plastic(**sy.PLASTIC_PARAMS)
ncA, loA, hiA, seedA = sy.c2_strain(0)
ncB, loB, hiB, seedB = sy.c2_strain(1)
epsA = mm.synth_uniform(N, ncA, seedA, loA, hiA, device=dev)
_, _, stA = mm.material_response(m, epsA, mm.initial_material_state(m, N, device=dev), tangent=False)
del epsA
epsB = mm.synth_uniform(N, ncB, seedB, loB, hiB, device=dev)


def buffers():
    return {"stress": torch.empty((6, N), dtype=torch.float64, device=dev), "tangent": torch.empty((36 if TANGENT else 1, N if TANGENT else 1), dtype=torch.float64, device=dev),
            "state": torch.empty((14, N), dtype=torch.float64, device=dev), "flags": torch.empty((N,), dtype=torch.uint8, device=dev),
            "iters": torch.empty((N,), dtype=torch.int16, device=dev)}


ref, out = buffers(), buffers()
opts = {"defer_check": True}
lib.mm_set_tuning(b"plastic_soa_prefetch", 0)
mm.material_response(m, epsB, stA, options=opts, tangent=TANGENT, out={k: v for k, v in ref.items() if TANGENT or k != "tangent"})
torch.cuda.synchronize()


def launch():
    return mm.material_response(m, epsB, stA, options=opts, tangent=TANGENT, out={k: v for k, v in out.items() if TANGENT or k != "tangent"})


def bits_differ():
    d = 0
    for k in ("stress", "tangent", "state"):
        d += int((out[k].view(torch.int64) != ref[k].view(torch.int64)).sum().item())
    for k in ("flags", "iters"):
        d += int((out[k] != ref[k]).sum().item())
    return d


iso = {v: [] for v in VARIANTS}
sus = {v: [] for v in VARIANTS}
diff = {}
for r in range(ROUNDS):
    for v in VARIANTS:
        lib.mm_set_tuning(b"plastic_soa_prefetch", v)
        for k in out.values():
            k.zero_()
        launch()
        torch.cuda.synchronize()
        if r == 0:
            diff[v] = bits_differ()
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            launch()
            b.record()
            torch.cuda.synchronize()
            iso[v].append(a.elapsed_time(b))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(SUST):
            launch()
        b.record()
        torch.cuda.synchronize()
        sus[v].append(a.elapsed_time(b) / SUST)
lib.mm_set_tuning(b"plastic_soa_prefetch", 0)
for v in VARIANTS:
    t = sorted(iso[v])
    print(json.dumps({"plastic_soa_prefetch": v, "points": N, "isolated_ms_median": round(t[len(t) // 2], 3), "isolated_ms_min": round(t[0], 3),
                      "back_to_back_ms": [round(x, 3) for x in sus[v]], "back_to_back_launches": SUST,
                      "tangent": TANGENT, "frac_of_6454_back_to_back": round((608.0 if TANGENT else 320.0) * N / (min(sus[v]) * 1e-3) / 1e9 / 6454.0, 4),
                      "words_differing_from_no_prefetch": diff[v]}), flush=True)
