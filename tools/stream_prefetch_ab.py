#!/usr/bin/env python3
"""A/B of the L2 prefetch hint of the streaming SoA kernels (mm_set_tuning "stream_soa_prefetch" = points ahead, 0 = no hint, -1 = the per-op defaults) on the other
BASELINE.json configs: LinearElastic with tangent, NeoHook / Yeoh from F with dS/dC, XuNeedleman 3D, 1e8 points each, alternating rounds
on one box.  Per distance: median of the isolated launches and the mean of a back-to-back run; outputs must equal the run without the hint.
usage: python tools/stream_prefetch_ab.py [points] [rounds] [back_to_back_launches] [distances,comma-separated]"""
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
SUST = int(sys.argv[3]) if len(sys.argv) > 3 else 40
DIST = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 16384, 32768, 65536, 131072]
lib = L.load()
dev = torch.device("cuda", 0)


def cases():
    nc, lo, hi, seed = sy.c1_strain()
    eps = mm.synth_uniform(N, nc, seed, lo, hi, device=dev)
    le = mm.LinearElastic(E=200e3, ν=0.3)
    yield "linear_elastic_with_tangent", 336, lambda: mm.material_response(le, eps)
    yield "linear_elastic_sigma_only", 96, lambda: mm.material_response(le, eps, tangent=False)
    del eps
    nc, lo, hi, seed = sy.c3_defgrad()
    F = mm.synth_uniform(N, nc, seed, lo, hi, device=dev)
    for name, prm in (("NeoHook", sy.NEOHOOK_PARAMS), ("Yeoh", sy.YEOH_PARAMS)):
        hm = getattr(mm, name)(λ=prm["lambda_"], μ=prm["mu"], c2=prm["c2"], c3=prm["c3"])
        yield name.lower() + "_from_F", 408, lambda hm=hm: mm.material_response(mm.dSdC(), hm, mm.DeformationGradient(F))
    del F
    x = sy.XU_PARAMS
    xm = mm.XuNeedleman(σₘₐₓ=x["sigma_max"], τₘₐₓ=x["tau_max"], Φₙ=mm.xu_needleman_Φₙ(x["sigma_max"], x["delta_n"]),
                        Φₜ=mm.xu_needleman_Φₜ(x["tau_max"], x["delta_t"]), Δₙˢ=x["Delta_n_star"])
    nc, lo, hi, seed = sy.c5_jump()
    D = mm.synth_uniform(N, nc, seed, lo, hi, device=dev)
    yield "xu_needleman_3d", 120, lambda: mm.material_response(xm, D)


def tensors(r):
    return [t for t in r if isinstance(t, torch.Tensor)]


for name, bpp, call in cases():
    lib.mm_set_tuning(b"stream_soa_prefetch", 0)
    ref = [t.clone() for t in tensors(call())]
    iso = {d: [] for d in DIST}
    sus = {d: [] for d in DIST}
    diff = {}
    for r in range(ROUNDS):
        for d in DIST:
            lib.mm_set_tuning(b"stream_soa_prefetch", d)
            got = tensors(call())
            torch.cuda.synchronize()
            if r == 0:
                diff[d] = sum(int((a.view(torch.int64) != b.view(torch.int64)).sum().item()) for a, b in zip(ref, got))
            del got
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                call()
                b.record()
                torch.cuda.synchronize()
                iso[d].append(a.elapsed_time(b))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(SUST):
                call()
            b.record()
            torch.cuda.synchronize()
            sus[d].append(a.elapsed_time(b) / SUST)
    lib.mm_set_tuning(b"stream_soa_prefetch", -1)
    for d in DIST:
        t = sorted(iso[d])
        print(json.dumps({"kernel": name, "stream_soa_prefetch": d, "points": N, "isolated_ms_median": round(t[len(t) // 2], 3), "isolated_ms_min": round(t[0], 3),
                          "back_to_back_ms": [round(x, 3) for x in sus[d]], "frac_of_6454_back_to_back": round(bpp * N / (min(sus[d]) * 1e-3) / 1e9 / 6454.0, 4),
                          "words_differing_from_no_hint": diff[d]}), flush=True)
    del ref
    torch.cuda.empty_cache()
