#!/usr/bin/env python3
"""Kernel timings for all five BASELINE.json configs (development aid; bench.py is the contract line).
Prints one JSON object per kernel: ms (median/min over reps), algorithmic GB/s, fraction of the measured HBM peak."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import materialmodels_jl_b200 as mm
from materialmodels_jl_b200 import synth as sy


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def report(name, N, bpp, med, mn, extra=None):
    gbs = bpp * N / (med * 1e-3) / 1e9
    d = {"kernel": name, "N": N, "ms_median": round(med, 4), "ms_min": round(mn, 4), "B_per_pt": bpp, "GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak(), 3),
         "updates_per_s": N / (med * 1e-3)}
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    want = lambda k: (not a.only) or any(w in k for w in a.only.split(","))
    S = a.scale
    if want("elastic"):
        m = mm.LinearElastic(E=200e3, ν=0.3)
        for N in (10**6, int(1e8 * S)):
            nc, lo, hi, seed = sy.c1_strain()
            eps = mm.synth_uniform(N, nc, seed, lo, hi)
            sig = torch.empty_like(eps)
            med, mn = timeit(lambda: mm.material_response(m, eps, tangent=False), a.reps)
            report("linear_elastic sigma-only soa", N, 96, med, mn)
            med, mn = timeit(lambda: mm.material_response(m, eps), a.reps)
            report("linear_elastic +tangent soa", N, 384, med, mn)
            del eps, sig
    if want("plastic"):
        N = int(1e8 * S)
        m = mm.Plastic(**sy.PLASTIC_PARAMS)
        for lname in ("soa", "aos"):
            ncA, loA, hiA, seedA = sy.c2_strain(0)
            ncB, loB, hiB, seedB = sy.c2_strain(1)
            eA = mm.synth_uniform(N, ncA, seedA, loA, hiA, layout=lname)
            _, _, stA = mm.material_response(m, eA, mm.initial_material_state(m, N, layout=lname), tangent=False)
            del eA
            eB = mm.synth_uniform(N, ncB, seedB, loB, hiB, layout=lname)
            shp = lambda nc: (nc, N) if lname == "soa" else (N, nc)
            out = {"stress": torch.empty(shp(6), dtype=torch.float64, device="cuda"), "tangent": torch.empty(shp(36), dtype=torch.float64, device="cuda"),
                   "state": torch.empty(shp(14), dtype=torch.float64, device="cuda"), "flags": torch.empty((N,), dtype=torch.uint8, device="cuda"),
                   "iters": torch.empty((N,), dtype=torch.int16, device="cuda")}
            med, mn = timeit(lambda: mm.material_response(m, eB, stA, options={"defer_check": True}, out=out), a.reps)
            frac = out["flags"].float().mean().item()
            report("plastic stepB " + lname, N, 608, med, mn, {"plastic_fraction": round(frac, 4)})
            if lname == "soa":
                z = mm.initial_material_state(m, N)
                e0 = mm.synth_uniform(N, 6, 1, -1e-4, 1e-4)
                med, mn = timeit(lambda: mm.material_response(m, e0, z, options={"defer_check": True}, out=out), a.reps)
                report("plastic all-elastic soa", N, 608, med, mn)
                e1 = mm.synth_uniform(N, 6, 2, 2e-3, 4e-3)
                e1[3:] = -e1[3:]
                med, mn = timeit(lambda: mm.material_response(m, e1, z, options={"defer_check": True}, out=out), a.reps)
                report("plastic all-plastic soa", N, 608, med, mn, {"plastic_fraction": out["flags"].float().mean().item()})
                del z, e0, e1
            del eB, stA, out
            torch.cuda.empty_cache()
    if want("hyper"):
        N = int(1e8 * S)
        nc, lo, hi, seed = sy.c3_defgrad()
        for lname in ("soa", "aos"):
            F = mm.synth_uniform(N, nc, seed, lo, hi, layout=lname)
            for name, prm in (("NeoHook", sy.NEOHOOK_PARAMS), ("Yeoh", sy.YEOH_PARAMS), ("StVenant", sy.NEOHOOK_PARAMS)):
                m = getattr(mm, name)(λ=prm["lambda_"], μ=prm["mu"], c2=prm["c2"], c3=prm["c3"])
                med, mn = timeit(lambda: mm.material_response(mm.dSdC(), m, mm.DeformationGradient(F), layout=lname), a.reps)
                report("%s from F %s" % (name, lname), N, 408, med, mn)
            del F
            torch.cuda.empty_cache()
    if want("traits"):
        N = int(1e8 * S)
        nc, lo, hi, seed = sy.c3_defgrad()
        F = mm.synth_uniform(N, nc, seed, lo, hi)
        m = mm.NeoHook(λ=1.0, μ=1.0)
        med, mn = timeit(lambda: mm.material_response(mm.dPdF(), m, mm.DeformationGradient(F)), a.reps)
        report("NeoHook dPdF from F soa", N, 8 * (9 + 9 + 81), med, mn)
        med, mn = timeit(lambda: mm.material_response(mm.dSdE(), m, mm.DeformationGradient(F)), a.reps)
        report("NeoHook dSdE from F soa", N, 408, med, mn)
        del F
        torch.cuda.empty_cache()
        F = mm.synth_uniform(N, nc, seed, lo, hi, layout="aos")
        med, mn = timeit(lambda: mm.material_response(mm.dPdF(), m, mm.DeformationGradient(F), layout="aos"), a.reps)
        report("NeoHook dPdF from F aos", N, 8 * (9 + 9 + 81), med, mn)
        del F
        torch.cuda.empty_cache()
    if want("transviso"):
        N = int(1e8 * S)
        m = mm.TransverselyIsotropic(E_L=100e3, E_T=10e3, G_LT=5e3, ν_LT=0.4, ν_TT=0.3)
        for lname in ("soa", "aos"):
            eps = mm.synth_uniform(N, 6, 99, -2e-3, 2e-3, layout=lname)
            a3 = mm.synth_uniform(N, 3, 100, 0.1, 1.0, layout=lname)
            a3 /= a3.norm(dim=0 if lname == "soa" else 1, keepdim=True)
            st = mm.TransverselyIsotropicState(a3.contiguous(), lname)
            med, mn = timeit(lambda: mm.material_response(m, eps, st), a.reps)
            report("transversely_isotropic " + lname, N, 8 * (6 + 3 + 6 + 36), med, mn)
            del eps, a3, st
            torch.cuda.empty_cache()
    if want("energy"):
        N = int(1e8 * S)
        nc, lo, hi, seed = sy.c3_defgrad()
        for lname in ("soa", "aos"):
            F = mm.synth_uniform(N, nc, seed, lo, hi, layout=lname)
            for name, prm in (("NeoHook", sy.NEOHOOK_PARAMS), ("Yeoh", sy.YEOH_PARAMS), ("StVenant", sy.NEOHOOK_PARAMS)):
                m = getattr(mm, name)(λ=prm["lambda_"], μ=prm["mu"], c2=prm["c2"], c3=prm["c3"])
                med, mn = timeit(lambda: mm.elastic_strain_energy_density(m, mm.DeformationGradient(F), layout=lname), a.reps)
                report("%s energy from F %s" % (name, lname), N, 80, med, mn)
            del F
            eps = mm.synth_uniform(N, 6, 5, -2e-3, 2e-3, layout=lname)
            le = mm.LinearElastic(E=200e3, ν=0.3)
            med, mn = timeit(lambda: mm.elastic_strain_energy_density(le, eps, layout=lname), a.reps)
            report("LinearElastic energy " + lname, N, 56, med, mn)
            del eps
            torch.cuda.empty_cache()
    if want("cells"):
        # the caller's quadrature loop fused around the J2 update.  Geometry is synthetic (timing only): ∇N and u are
        # uniform random, scaled so that about half the points yield.  Algorithmic bytes per POINT: dNdx 24 nn + dΩ 8 + state 112 in,
        # state 112 + 2 flag bytes out, per cell ue and fe 24 nn each (divided by nq); + 288 when the tangent is written.
        m = mm.Plastic(**sy.PLASTIC_PARAMS)
        for nn, nq in ((4, 1), (8, 8)):
            P = int(1e8 * S) // nq * nq
            ncell = P // nq
            amp, best = 1e-4, (2.0, 1e-4)                      # displacement amplitude giving ~50 % plastic points (calibrated on 2^20 points)
            Pc = (1 << 20) // nq * nq
            dNc, dOc = mm.synth_uniform(Pc, 3 * nn, 71, -1.0, 1.0), mm.synth_uniform(Pc, 1, 72, 0.5, 1.5).reshape(Pc)
            for _ in range(24):
                _, stc, _ = mm.cell_internal_forces(m, dNc, dOc, mm.synth_uniform(Pc // nq, 3 * nn, 73, -amp, amp), mm.initial_material_state(m, Pc), options={"defer_check": True})
                fr = (stc.flags & 1).float().mean().item()
                if abs(fr - 0.5) < best[0]:
                    best = (abs(fr - 0.5), amp)
                amp *= 1.12
            amp = best[1]
            del dNc, dOc, stc
            dN = mm.synth_uniform(P, 3 * nn, 71, -1.0, 1.0)
            dO = mm.synth_uniform(P, 1, 72, 0.5, 1.5).reshape(P)
            ue = mm.synth_uniform(ncell, 3 * nn, 73, -amp, amp)
            _, stA, _ = mm.cell_internal_forces(m, dN, dO, ue, mm.initial_material_state(m, P), options={"defer_check": True})
            ue = mm.synth_uniform(ncell, 3 * nn, 74, -amp, amp)
            for tan in (False, True):
                med, mn = timeit(lambda: mm.cell_internal_forces(m, dN, dO, ue, stA, tangent=tan, options={"defer_check": True}), a.reps)
                _, stB, _ = mm.cell_internal_forces(m, dN, dO, ue, stA, tangent=tan, options={"defer_check": True})
                bpp = 24 * nn + 8 + 112 + 112 + 2 + 48 * nn / nq + (288 if tan else 0)
                report("cells J2 nn=%d nq=%d %s" % (nn, nq, "fe+state+tangent" if tan else "fe+state"), P, bpp, med, mn,
                       {"plastic_fraction": round((stB.flags & 1).float().mean().item(), 3), "n_fail": int(stB.n_fail.item()), "unfused_B_per_pt": 608 + 48 + 24 * nn + 8 + 48 * nn / nq})
            # element stiffness BᵀDB as well (D stays in registers): + 8 (3nn)²/nq bytes per point; fewer points so that ke fits
            Pk = min(P, int(2e7 * S) // nq * nq if nq == 1 else P)
            nck = Pk // nq
            ke_bytes = 8 * 9 * nn * nn / nq
            sub = lambda x, n: x[..., :n].contiguous()
            dNk, dOk, uek, stk = sub(dN, Pk), dO[:Pk].contiguous(), sub(ue, nck), mm.PlasticState(sub(stA.data, Pk))
            med, mn = timeit(lambda: mm.cell_internal_forces(m, dNk, dOk, uek, stk, stiffness=True, options={"defer_check": True}), a.reps)
            flop = 2 * (nn * 54 + nn * nn * 27)
            report("cells J2 nn=%d nq=%d fe+state+ke" % (nn, nq), Pk, 24 * nn + 8 + 112 + 112 + 2 + 48 * nn / nq + ke_bytes, med, mn,
                   {"ke_flop_per_pt": flop, "ke_TFLOPs": round(flop * Pk / (med * 1e-3) / 1e12, 2)})
            del dN, dO, ue, stA, stB, dNk, dOk, uek, stk
            torch.cuda.empty_cache()
    if want("hypercells"):
        hm = mm.NeoHook(λ=1.0, μ=1.0)
        for nn, nq in ((4, 1), (8, 8)):
            P = int(1e8 * S) // nq * nq
            ncell = P // nq
            dN = mm.synth_uniform(P, 3 * nn, 71, -1.0, 1.0)
            dO = mm.synth_uniform(P, 1, 72, 0.5, 1.5).reshape(P)
            ue = mm.synth_uniform(ncell, 3 * nn, 74, -0.03, 0.03)
            for tan in (False, True):
                med, mn = timeit(lambda: mm.cell_internal_forces(hm, dN, dO, ue, tangent=tan), a.reps)
                report("cells NeoHook nn=%d nq=%d %s" % (nn, nq, "fe+dPdF" if tan else "fe"), P, 24 * nn + 8 + 48 * nn / nq + (648 if tan else 0), med, mn)
            del dN, dO, ue
            torch.cuda.empty_cache()
    if want("xu"):
        N = int(1e8 * S)
        x = sy.XU_PARAMS
        m = mm.XuNeedleman(σₘₐₓ=x["sigma_max"], τₘₐₓ=x["tau_max"], Φₙ=mm.xu_needleman_Φₙ(x["sigma_max"], x["delta_n"]),
                           Φₜ=mm.xu_needleman_Φₜ(x["tau_max"], x["delta_t"]), Δₙˢ=x["Delta_n_star"])
        nc, lo, hi, seed = sy.c5_jump()
        for lname in ("soa", "aos"):
            D = mm.synth_uniform(N, nc, seed, lo, hi, layout=lname)
            med, mn = timeit(lambda: mm.material_response(m, D, layout=lname), a.reps)
            report("xu_needleman 3D " + lname, N, 120, med, mn)
            del D
    if want("wrapped"):
        N = int(1e8 * S)
        m = mm.Plastic(**sy.PLASTIC_PARAMS)
        for dim, nc, amp in ((mm.PlaneStress(), 3, 1.2e-3), (mm.UniaxialStress(), 1, 2.5e-3), (mm.ShellStress(), 6, 0.8e-3), (mm.PlaneStrain(), 3, 1.0e-3)):
            eA = mm.synth_uniform(N, nc, 31, -amp, amp)
            _, _, stA = mm.material_response(dim, m, eA, mm.initial_material_state(m, N), options={"defer_check": True})
            eB = mm.synth_uniform(N, nc, 32, -amp, amp)
            med, mn = timeit(lambda: mm.material_response(dim, m, eB, stA, options={"defer_check": True}), a.reps)
            _, _, stB = mm.material_response(dim, m, eB, stA, options={"defer_check": True})
            bpp = 8 * (nc + 14 + nc + nc * nc + 14) + 3
            report("plastic %s stepB soa" % type(dim).__name__, N, bpp, med, mn,
                   {"plastic_fraction": round((stB.flags & 1).float().mean().item(), 3), "mean_outer_iters": round(stB.outer_iters.float().mean().item(), 2),
                    "n_fail": int(stB.n_fail.item())})
            del eA, eB, stA, stB
            torch.cuda.empty_cache()
        dim, nc, amp = mm.PlaneStress(), 3, 1.2e-3                       # the Julia zero-copy layout through the TMA tile kernel
        eA = mm.synth_uniform(N, nc, 31, -amp, amp, layout="aos")
        _, _, stA = mm.material_response(dim, m, eA, mm.initial_material_state(m, N, layout="aos"), options={"defer_check": True})
        eB = mm.synth_uniform(N, nc, 32, -amp, amp, layout="aos")
        med, mn = timeit(lambda: mm.material_response(dim, m, eB, stA, options={"defer_check": True}), a.reps)
        report("plastic PlaneStress stepB aos", N, 8 * (nc + 14 + nc + nc * nc + 14) + 3, med, mn)
        del eA, eB, stA
        torch.cuda.empty_cache()
    if want("wrapti"):
        N = int(1e8 * S)
        m = mm.TransverselyIsotropic(E_L=100e3, E_T=10e3, G_LT=5e3, ν_LT=0.4, ν_TT=0.3)
        a3 = mm.synth_uniform(N, 3, 100, 0.1, 1.0)
        a3 /= a3.norm(dim=0, keepdim=True)
        st = mm.TransverselyIsotropicState(a3.contiguous(), "soa")
        for dim, nc in ((mm.PlaneStress(), 3), (mm.ShellStress(), 6)):
            e = mm.synth_uniform(N, nc, 41, -1e-3, 1e-3)
            med, mn = timeit(lambda: mm.material_response(dim, m, e, st, options={"defer_check": True}), a.reps)
            report("transversely_isotropic %s soa" % type(dim).__name__, N, 8 * (nc + 3 + nc + nc * nc) + 1, med, mn)
            del e
        del a3, st
        torch.cuda.empty_cache()
    if want("wrapxtal"):
        N = int(2e6 * S)
        m = mm.CrystalViscoPlastic(slipsystems=mm.slipsystems(mm.FCC(), sy.CRYSTAL_RODRIGUES), **sy.CRYSTAL_PARAMS)
        for dim, nc, amp in ((mm.PlaneStress(), 3, 4e-3), (mm.UniaxialStress(), 1, 6e-3)):
            e = mm.synth_uniform(N, nc, 51, -amp, amp)
            st0 = mm.initial_material_state(m, N)
            med, mn = timeit(lambda: mm.material_response(dim, m, e, st0, 1.0, options={"defer_check": True}), max(a.reps // 3, 2), warm=1)
            _, _, st1 = mm.material_response(dim, m, e, st0, 1.0, options={"defer_check": True})
            report("crystal %s soa" % type(dim).__name__, N, 8 * (nc + 42 + nc + nc * nc + 42) + 4, med, mn,
                   {"active_fraction": round(((st1.flags & 0xFFF) != 0).float().mean().item(), 3), "mean_outer_iters": round(st1.outer_iters.float().mean().item(), 2),
                    "n_fail": int(st1.n_fail.item())})
            del e, st0, st1
            torch.cuda.empty_cache()
    if want("xtalvariants"):
        # crystal plasticity off the headline configuration: BCC table (the reference's, with two non-slip 'systems' -> fix-up points)
        # and cross hardening q != 0 (Sherman-Morrison path)
        N = int(1e7 * S)
        for tag, structure, q in (("FCC q=0.3", mm.FCC(), 0.3), ("BCC q=0", mm.BCC(), 0.0), ("BCC q=0.3", mm.BCC(), 0.3)):
            prm = dict(sy.CRYSTAL_PARAMS); prm["q"] = q
            m = mm.CrystalViscoPlastic(slipsystems=mm.slipsystems(structure, sy.CRYSTAL_RODRIGUES), **prm)
            nc, lo, hi, seed = sy.c4_strain_increment(0)
            dA = mm.synth_uniform(N, nc, seed, lo, hi)
            _, _, stA = mm.material_response(m, dA, mm.initial_material_state(m, N), 1.0, options={"defer_check": True})
            nc, lo, hi, seed = sy.c4_strain_increment(1)
            dB = mm.synth_uniform(N, nc, seed, lo, hi)
            med, mn = timeit(lambda: mm.material_response(m, dB, stA, 1.0, options={"defer_check": True}), max(a.reps // 3, 2), warm=1)
            _, _, stB = mm.material_response(m, dB, stA, 1.0, options={"defer_check": True})
            report("crystal stepB " + tag, N, 1056, med, mn, {"mean_iters": stB.iters.float().mean().item(), "n_fail": int(stB.n_fail.item())})
            del dA, dB, stA, stB
            torch.cuda.empty_cache()
    if want("crystal"):
        N = int(1e7 * S)
        ss = mm.slipsystems(mm.FCC(), sy.CRYSTAL_RODRIGUES)
        m = mm.CrystalViscoPlastic(slipsystems=ss, **sy.CRYSTAL_PARAMS)
        for lname in ("soa", "aos"):
            nc, lo, hi, seed = sy.c4_strain_increment(0)
            dA = mm.synth_uniform(N, nc, seed, lo, hi, layout=lname)
            z = mm.initial_material_state(m, N, layout=lname)
            med, mn = timeit(lambda: mm.material_response(m, dA, z, 1.0, options={"defer_check": True}), max(a.reps // 3, 2), warm=1)
            _, _, stA = mm.material_response(m, dA, z, 1.0)
            report("crystal stepA " + lname, N, 1056, med, mn, {"mean_iters": stA.iters.float().mean().item(), "frac_active": (stA.flags != 0).float().mean().item()})
            nc, lo, hi, seed = sy.c4_strain_increment(1)
            dB = mm.synth_uniform(N, nc, seed, lo, hi, layout=lname)
            med, mn = timeit(lambda: mm.material_response(m, dB, stA, 1.0, options={"defer_check": True}), max(a.reps // 3, 2), warm=1)
            _, _, stB = mm.material_response(m, dB, stA, 1.0, options={"defer_check": True})
            report("crystal stepB " + lname, N, 1056, med, mn, {"mean_iters": stB.iters.float().mean().item(), "frac_active": (stB.flags != 0).float().mean().item(),
                                                                 "n_fail": int(stB.n_fail.item())})


if __name__ == "__main__":
    main()
