#!/usr/bin/env python3
"""Is the bimodal XuNeedleman SoA timing (0.84 or 1.00 of the HBM peak, by box / run) a matter of where the component rows land?
Times the device-resident call for a few point counts around 1e8 (the row stride of every SoA array is 8 N bytes), each with fresh allocations,
several times per process.  usage: xu_stride_probe.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import materialmodels_jl_b200 as mm
from materialmodels_jl_b200 import synth as sy

x = sy.XU_PARAMS
m = mm.XuNeedleman(σₘₐₓ=x["sigma_max"], τₘₐₓ=x["tau_max"], Φₙ=mm.xu_needleman_Φₙ(x["sigma_max"], x["delta_n"]),
                   Φₜ=mm.xu_needleman_Φₜ(x["tau_max"], x["delta_t"]), Δₙˢ=x["Delta_n_star"])
nc, lo, hi, seed = sy.c5_jump()
for N in (10**8, 10**8 + 1024, 10**8 - 256 * 1000, 99_999_744 + 2**16, 2**27 // 8 * 6):
    for trial in range(3):
        D = mm.synth_uniform(N, nc, seed, lo, hi)
        out = {"stress": torch.empty((3, N), dtype=torch.float64, device="cuda"), "tangent": torch.empty((9, N), dtype=torch.float64, device="cuda")}
        for _ in range(3):
            mm.material_response(m, D, out=out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); mm.material_response(m, D, out=out); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        print(json.dumps({"N": N, "trial": trial, "ms_median": round(ts[5], 4), "ms_min": round(ts[0], 4), "GBps": round(120 * N / ts[5] / 1e6, 1),
                          "D_ptr_mod_2MiB": D.data_ptr() % (1 << 21), "T_ptr_mod_2MiB": out["stress"].data_ptr() % (1 << 21)}), flush=True)
        del D, out
        torch.cuda.empty_cache()
