"""Trait layer (reference src/traits.jl:34-107): strain kinds C / F / E in, tangent kinds
∂S∂C / ∂S∂E / ∂Pᵀ∂F out.  Oracle vs the reference's own checks (test/test_straintraits.jl), kernels' math on the host vs
the oracle, CUDA path vs the oracle."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
import materialmodels_jl_b200._abi as abi
import materialmodels_jl_b200.synth as sy

MODELS = [("NeoHook", abi.MM_NEOHOOK, sy.NEOHOOK_PARAMS), ("Yeoh", abi.MM_YEOH, sy.YEOH_PARAMS), ("StVenant", abi.MM_STVENANT, sy.NEOHOOK_PARAMS)]
SLOTS = ((0, 0), (1, 0), (2, 0), (1, 1), (2, 1), (2, 2))


def C_and_E_from_F(F):   # F: (9, N) column-major per point
    Fm = F.reshape(3, 3, -1, order="F")
    Cm = np.einsum("kin,kjn->ijn", Fm, Fm)
    C6 = np.stack([Cm[i, j] for i, j in SLOTS])
    E6 = C6.copy()
    E6[[0, 3, 5]] -= 1.0
    return np.ascontiguousarray(C6), np.ascontiguousarray(E6 / 2)


def test_oracle_reproduces_test_straintraits(oracle):
    F = np.array([2.0, 0, 0, 1, 1, 0, 0, 0, 1]).reshape(9, 1)                       # test_straintraits.jl:5
    prm = (1.0, 1.0, 0.0, 0.0)
    S, T = oracle.hyper_ex(abi.MM_NEOHOOK, prm, F, oracle.STRAIN_F, oracle.TAN_DSDC)
    Pt, dP = oracle.hyper_ex(abi.MM_NEOHOOK, prm, F, oracle.STRAIN_F, oracle.TAN_DPDF)
    Fm = F[:, 0].reshape(3, 3, order="F")
    Sm = np.array([[S[0, 0], S[1, 0], S[2, 0]], [S[1, 0], S[3, 0], S[4, 0]], [S[2, 0], S[4, 0], S[5, 0]]])
    assert np.array_equal(Pt[:, 0].reshape(3, 3, order="F"), Sm @ Fm.T)             # :27  Pᵀ == S⋅F'
    SE, TE = oracle.hyper_ex(abi.MM_NEOHOOK, prm, F, oracle.STRAIN_F, oracle.TAN_DSDE)
    assert np.array_equal(SE, S) and np.array_equal(2 * T, TE)                       # :30-31
    Pa, da = oracle.neohook_dPdF_ad(prm, F[:, 0])
    assert np.allclose(dP[:, 0], da, rtol=1e-13, atol=1e-13) and np.allclose(Pt[:, 0], Pa, rtol=1e-14)   # :41-48 dPᵀdF ≈ AD
    C6, E6 = C_and_E_from_F(F)
    S2, T2 = oracle.hyper_ex(abi.MM_NEOHOOK, prm, E6, oracle.STRAIN_E, oracle.TAN_DSDE)   # :12-18 strain transformations
    assert np.allclose(S2, S, rtol=1e-15) and np.allclose(T2, TE, rtol=1e-15)
    rng = np.random.default_rng(5)
    for _ in range(10):                                                              # AD identity on random well-conditioned F
        F = (np.eye(3) + 0.2 * (2 * rng.random((3, 3)) - 1)).reshape(9, 1, order="F")
        Pt, dP = oracle.hyper_ex(abi.MM_NEOHOOK, prm, F, oracle.STRAIN_F, oracle.TAN_DPDF)
        Pa, da = oracle.neohook_dPdF_ad(prm, F[:, 0])
        assert np.linalg.norm(dP[:, 0] - da) < 1e-13 * np.linalg.norm(da)


@pytest.mark.parametrize("name,model,prm", MODELS)
def test_twin_traits_match_oracle(oracle, name, model, prm):
    tw = H.hosttwin()
    N = 5000
    nc, lo, hi, seed = sy.c3_defgrad()
    F = sy.host_uniform(N, nc, seed, lo, hi)
    C6, E6 = C_and_E_from_F(F)
    hp = abi.MMHyper(**prm)
    for strain, kind in ((F, 1), (C6, 0), (E6, 2)):
        for tan in (0, 1, 2):
            if tan == 2 and kind != 1:
                continue
            S_ref, T_ref = oracle.hyper_ex(model, hp, strain, kind, tan)
            S, T = np.zeros_like(S_ref), np.zeros_like(T_ref)
            assert tw.twin_hyper_ex(model, C.byref(hp), H.P(strain), kind, tan, H.P(S), H.P(T), C.c_int64(N)) == 0
            assert H.rel_err(S, S_ref).max() < H.RTOL and H.rel_err(T, T_ref).max() < H.RTOL, (kind, tan)


@pytest.mark.gpu
@pytest.mark.parametrize("lname,lay", [("soa", 0), ("aos", 1)])
@pytest.mark.parametrize("name,model,prm", MODELS)
def test_gpu_traits_match_oracle(mm, cuda_lib, oracle, name, model, prm, lname, lay):
    import torch

    N = 60007
    m = getattr(mm, name)(λ=prm["lambda_"], μ=prm["mu"], c2=prm["c2"], c3=prm["c3"])
    nc, lo, hi, seed = sy.c3_defgrad()
    F = sy.host_uniform(N, nc, seed, lo, hi)
    C6, E6 = C_and_E_from_F(F)
    hp = abi.MMHyper(**prm)
    lay_of = (lambda a: a) if lay == 0 else (lambda a: np.ascontiguousarray(a.T))
    soa = (lambda a: a) if lay == 0 else (lambda a: np.ascontiguousarray(a.T))
    routes = [(mm.dSdC(), mm.DeformationGradient, F, 1, 0), (mm.dSdE(), mm.DeformationGradient, F, 1, 1), (mm.dPdF(), mm.DeformationGradient, F, 1, 2),
              (mm.dSdE(), mm.GreenLagrange, E6, 2, 1), (mm.dSdC(), mm.GreenLagrange, E6, 2, 0), (mm.dSdE(), mm.RightCauchyGreen, C6, 0, 1)]
    for tag, wrap, strain, kind, tan in routes:
        S_ref, T_ref = oracle.hyper_ex(model, hp, strain, kind, tan)
        S, T, _ = mm.material_response(tag, m, wrap(torch.from_numpy(lay_of(strain)).cuda()), None, 0.0, layout=lname)
        assert H.rel_err(soa(S.cpu().numpy()), S_ref).max() < H.RTOL, (kind, tan)
        assert H.rel_err(soa(T.cpu().numpy()), T_ref).max() < H.RTOL, (kind, tan)
    # test_straintraits.jl:27,30-31 through the CUDA path
    S, T, _ = mm.material_response(mm.dSdC(), m, mm.DeformationGradient(torch.from_numpy(lay_of(F)).cuda()), layout=lname)
    SE, TE, _ = mm.material_response(mm.dSdE(), m, mm.DeformationGradient(torch.from_numpy(lay_of(F)).cuda()), layout=lname)
    assert torch.equal(S, SE) and torch.equal(2 * T, TE)
