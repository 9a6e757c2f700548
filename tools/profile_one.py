#!/usr/bin/env python3
"""Tiny driver for ncu captures: runs a few launches of ONE kernel at a size ncu can replay quickly.
usage: profile_one.py {plastic|plastic_aos|crystal|neohook|xu|elastic|wrapped_planestress|...} [points]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import materialmodels_jl_b200 as mm
from materialmodels_jl_b200 import synth as sy

which = sys.argv[1] if len(sys.argv) > 1 else "plastic"
N = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10**7
reps = 3
if which.startswith("plastic"):
    lname = "aos" if which.endswith("aos") else "soa"
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    nc, lo, hi, seed = sy.c2_strain(0)
    eA = mm.synth_uniform(N, nc, seed, lo, hi, layout=lname)
    _, _, stA = mm.material_response(m, eA, mm.initial_material_state(m, N, layout=lname), tangent=False)
    nc, lo, hi, seed = sy.c2_strain(1)
    eB = mm.synth_uniform(N, nc, seed, lo, hi, layout=lname)
    for _ in range(reps):
        s, t, st = mm.material_response(m, eB, stA, options={"defer_check": True})
elif which in ("crystal", "crystal_aos"):
    lname = "aos" if which.endswith("aos") else "soa"
    ss = mm.slipsystems(mm.FCC(), sy.CRYSTAL_RODRIGUES)
    m = mm.CrystalViscoPlastic(slipsystems=ss, **sy.CRYSTAL_PARAMS)
    nc, lo, hi, seed = sy.c4_strain_increment(0)
    dA = mm.synth_uniform(N, nc, seed, lo, hi, layout=lname)
    _, _, stA = mm.material_response(m, dA, mm.initial_material_state(m, N, layout=lname), 1.0)
    nc, lo, hi, seed = sy.c4_strain_increment(1)
    dB = mm.synth_uniform(N, nc, seed, lo, hi, layout=lname)
    for _ in range(reps):
        mm.material_response(m, dB, stA, 1.0, options={"defer_check": True})
elif which == "neohook":
    nc, lo, hi, seed = sy.c3_defgrad()
    F = mm.synth_uniform(N, nc, seed, lo, hi)
    m = mm.NeoHook(λ=1.0, μ=1.0)
    for _ in range(reps):
        mm.material_response(mm.dSdC(), m, mm.DeformationGradient(F))
elif which == "xu":
    x = sy.XU_PARAMS
    m = mm.XuNeedleman(σₘₐₓ=x["sigma_max"], τₘₐₓ=x["tau_max"], Φₙ=mm.xu_needleman_Φₙ(x["sigma_max"], x["delta_n"]),
                       Φₜ=mm.xu_needleman_Φₜ(x["tau_max"], x["delta_t"]), Δₙˢ=x["Delta_n_star"])
    nc, lo, hi, seed = sy.c5_jump()
    D = mm.synth_uniform(N, nc, seed, lo, hi)
    for _ in range(reps):
        mm.material_response(m, D)
elif which in ("cells_tet4", "cells_hex8", "cells_hex8_ke"):
    nn, nq = (4, 1) if which.endswith("tet4") else (8, 8)
    P = N // nq * nq
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    amp = 7e-4 if nq == 1 else 3.6e-4
    dN = mm.synth_uniform(P, 3 * nn, 71, -1.0, 1.0)
    dO = mm.synth_uniform(P, 1, 72, 0.5, 1.5).reshape(P)
    _, stA, _ = mm.cell_internal_forces(m, dN, dO, mm.synth_uniform(P // nq, 3 * nn, 73, -amp, amp), mm.initial_material_state(m, P), options={"defer_check": True})
    ue = mm.synth_uniform(P // nq, 3 * nn, 74, -amp, amp)
    for _ in range(reps):
        mm.cell_internal_forces(m, dN, dO, ue, stA, stiffness=which.endswith("_ke"), options={"defer_check": True})
elif which == "transviso":
    m = mm.TransverselyIsotropic(E_L=100e3, E_T=10e3, G_LT=5e3, ν_LT=0.4, ν_TT=0.3)
    eps = mm.synth_uniform(N, 6, 99, -2e-3, 2e-3)
    a3 = mm.synth_uniform(N, 3, 100, 0.1, 1.0)
    a3 /= a3.norm(dim=0, keepdim=True)
    st = mm.TransverselyIsotropicState(a3.contiguous(), "soa")
    for _ in range(reps):
        mm.material_response(m, eps, st)
elif which.startswith("wrapped_"):          # wrapped_planestress / wrapped_uniaxialstress / wrapped_shellstress (J2 inside)
    dim, nc, amp = {"planestress": (mm.PlaneStress(), 3, 1.2e-3), "uniaxialstress": (mm.UniaxialStress(), 1, 2.5e-3),
                    "shellstress": (mm.ShellStress(), 6, 0.8e-3), "planestrain": (mm.PlaneStrain(), 3, 1.0e-3)}[which.split("_", 1)[1]]
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    eA = mm.synth_uniform(N, nc, 31, -amp, amp)
    _, _, stA = mm.material_response(dim, m, eA, mm.initial_material_state(m, N), options={"defer_check": True})
    eB = mm.synth_uniform(N, nc, 32, -amp, amp)
    for _ in range(reps):
        mm.material_response(dim, m, eB, stA, options={"defer_check": True})
elif which.startswith("wrapxtal_"):         # wrapxtal_planestress / wrapxtal_uniaxialstress: the rounds of the two-kernel crystal path (mm_wrap.cu)
    dim, nc, amp = {"planestress": (mm.PlaneStress(), 3, 4e-3), "uniaxialstress": (mm.UniaxialStress(), 1, 6e-3)}[which.split("_", 1)[1]]
    m = mm.CrystalViscoPlastic(slipsystems=mm.slipsystems(mm.FCC(), sy.CRYSTAL_RODRIGUES), **sy.CRYSTAL_PARAMS)
    e = mm.synth_uniform(N, nc, 51, -amp, amp)
    st0 = mm.initial_material_state(m, N)
    for _ in range(2):
        mm.material_response(dim, m, e, st0, 1.0, options={"defer_check": True})
elif which == "elastic":
    nc, lo, hi, seed = sy.c1_strain()
    e = mm.synth_uniform(N, nc, seed, lo, hi)
    m = mm.LinearElastic(E=200e3, ν=0.3)
    for _ in range(reps):
        mm.material_response(m, e)
torch.cuda.synchronize()
print("done", which, N)
