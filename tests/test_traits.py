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


# ---- small-strain trait route and Dim{d} (reference src/traits.jl:91-148, src/wrappers.jl:3,29-41; test_linear_elastic.jl:32-33) ----
def test_native_types_and_small_strain_errors_without_gpu(mm):
    assert mm.native_strain_type(mm.LinearElastic) is mm.SmallStrain                    # test_linear_elastic.jl:32
    assert mm.native_stress_type(mm.LinearElastic) == "TrueStress"                      # test_linear_elastic.jl:33
    assert mm.native_strain_type(mm.NeoHook) is mm.RightCauchyGreen and mm.native_stress_type(mm.Yeoh) == "SecondPiolaKirchhoff"
    m = mm.LinearElastic(E=200e3, ν=0.3)
    eps = np.zeros((6, 3))
    with pytest.raises(RuntimeError, match="DeformationGradient is not the native strain measure for LinearElastic:"):   # traits.jl:120
        mm.material_response(mm.dsigmadeps(), m, mm.DeformationGradient(np.zeros((9, 3))))
    with pytest.raises(TypeError, match="no transformation"):                           # no transform_stress_tangent method (traits.jl:58-77)
        mm.material_response(mm.dSdC(), m, mm.SmallStrain(eps))
    with pytest.raises(TypeError, match="no transformation"):
        mm.material_response(mm.dsigmadeps(), mm.NeoHook(λ=1.0, μ=1.0), mm.RightCauchyGreen(np.zeros((6, 3))))
    with pytest.raises(ValueError):
        mm.Dim(4)
    assert mm.Dim(2).KIND == mm.PlaneStrain.KIND and mm.Dim(1).KIND == mm.UniaxialStrain.KIND and mm.Dim(3).KIND == 0


@pytest.mark.gpu
def test_small_strain_trait_route_is_the_native_call(mm, cuda_lib):
    """material_response(∂σ∂ε(), m, SmallStrain(ε), state, Δt) == material_response(m, ε, state, Δt) for every small-strain material"""
    import torch

    N = 5000
    eq = lambda a, b: torch.equal(a, b)
    le = mm.LinearElastic(E=200e3, ν=0.3)
    nc, lo, hi, seed = sy.c1_strain()
    eps = mm.synth_uniform(N, nc, seed, lo, hi)
    a = mm.material_response(le, eps)
    b = mm.material_response(mm.dsigmadeps(), le, mm.SmallStrain(eps))
    assert eq(a[0], b[0]) and eq(a[1], b[1])
    pl = mm.Plastic(**sy.PLASTIC_PARAMS)
    nc, lo, hi, seed = sy.c2_strain(0)
    e2 = mm.synth_uniform(N, nc, seed, lo, hi)
    st = mm.initial_material_state(pl, N)
    a = mm.material_response(pl, e2, st)
    b = mm.material_response(mm.dsigmadeps(), pl, mm.SmallStrain(e2), st, 0.0, cache=mm.get_cache(pl))
    assert eq(a[0], b[0]) and eq(a[1], b[1]) and eq(a[2].data, b[2].data) and eq(a[2].iters, b[2].iters)
    ti = mm.TransverselyIsotropic(E_L=140e3, E_T=10e3, G_LT=5e3, ν_LT=0.3, ν_TT=0.4)
    ts = mm.initial_material_state(ti, (0.0, 0.6, 0.8), N)
    a = mm.material_response(ti, eps, ts)
    b = mm.material_response(mm.dsigmadeps(), ti, mm.SmallStrain(eps), ts)
    assert eq(a[0], b[0]) and eq(a[1], b[1])
    cm = mm.CrystalViscoPlastic(slipsystems=mm.slipsystems(mm.FCC(), sy.CRYSTAL_RODRIGUES), **sy.CRYSTAL_PARAMS)
    nc, lo, hi, seed = sy.c4_strain_increment(0)
    de = mm.synth_uniform(N, nc, seed, lo, hi)
    cs = mm.initial_material_state(cm, N)
    a = mm.material_response(cm, de, cs, 1.0)
    b = mm.material_response(mm.dsigmadeps(), cm, mm.SmallStrain(de), cs, 1.0)
    assert eq(a[0], b[0]) and eq(a[1], b[1]) and eq(a[2].data, b[2].data)
    # numpy strains take the host-buffer entry points through the same route
    s_h, t_h, _ = mm.material_response(mm.dsigmadeps(), le, mm.SmallStrain(eps.cpu().numpy()))
    assert np.array_equal(s_h, a_np(mm.material_response(le, eps)[0]))


def a_np(t):
    return t.cpu().numpy()


@pytest.mark.gpu
def test_generic_dim_wrapper(mm, cuda_lib):
    """Dim{d} (wrappers.jl:3): Dim(2) is dispatched with PlaneStrain, Dim(1) with UniaxialStrain (:29-41); Dim(3) is the 3D call (:22)"""
    import torch

    N = 4000
    pl = mm.Plastic(**sy.PLASTIC_PARAMS)
    st = mm.initial_material_state(pl, N)
    for d, ref_dim, nc in ((2, mm.PlaneStrain(), 3), (1, mm.UniaxialStrain(), 1)):
        e = mm.synth_uniform(N, nc, 77 + d, -1.5e-3, 1.5e-3)
        a = mm.material_response(ref_dim, pl, e, st)
        b = mm.material_response(mm.Dim(d), pl, e, st)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2].data, b[2].data)
    nc, lo, hi, seed = sy.c2_strain(0)
    e3 = mm.synth_uniform(N, nc, seed, lo, hi)
    a = mm.material_response(pl, e3, st)
    b = mm.material_response(mm.Dim(3), pl, e3, st)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2].data, b[2].data)
