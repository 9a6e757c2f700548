"""TransverselyIsotropic (reference src/transverselyisotropic.jl:19-106) and batched
elastic_strain_energy_density (LinearElastic.jl:63, neohook.jl:29, yeoh.jl:31, stvenant.jl:28, traits.jl:133-136; rank 4).
Oracle vs the reference's own checks (test/test_transversely.jl, test_neohook.jl:5-6, test_yeoh.jl:6-7,
test_straintraits.jl:51-54), kernels' math on the host vs the oracle, CUDA path vs the oracle."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
import materialmodels_jl_b200._abi as abi
import materialmodels_jl_b200.synth as sy

TI = dict(E_L=100e3, E_T=10e3, G_LT=5e3, nu_LT=0.4, nu_TT=0.3)          # test_transversely.jl:4-8
SLOTS = ((0, 0), (1, 0), (2, 0), (1, 1), (2, 1), (2, 2))
HYPER = [("NeoHook", abi.MM_NEOHOOK, sy.NEOHOOK_PARAMS), ("Yeoh", abi.MM_YEOH, sy.YEOH_PARAMS), ("StVenant", abi.MM_STVENANT, sy.NEOHOOK_PARAMS)]


def ti_params(oracle):
    return oracle.transviso_params(TI["E_L"], TI["E_T"], TI["G_LT"], TI["nu_LT"], TI["nu_TT"])


def full4(t36):
    """36-component minor-symmetric storage -> full 3x3x3x3"""
    E = np.zeros((3, 3, 3, 3))
    for I, (i, j) in enumerate(SLOTS):
        for J, (k, l) in enumerate(SLOTS):
            E[i, j, k, l] = E[j, i, k, l] = E[i, j, l, k] = E[j, i, l, k] = t36[I + 6 * J]
    return E


def sym6(m):
    return np.array([m[i, j] for i, j in SLOTS])


def unit_directions(N, seed):
    a = sy.host_uniform(N, 3, seed, [-1.0] * 3, [1.0] * 3)
    a[:, np.linalg.norm(a, axis=0) < 1e-3] = np.array([[1.0], [0.0], [0.0]])
    return np.ascontiguousarray(a / np.linalg.norm(a, axis=0))


def test_oracle_reproduces_test_transversely(oracle):
    p = ti_params(oracle)
    rng = np.random.default_rng(7)
    e = rng.random((3, 3))
    e = 0.5 * (e + e.T)
    eps = sym6(e).reshape(6, 1)
    a = np.array([[1.0], [0.0], [0.0]])
    sig, tan = oracle.transversely_isotropic(p, eps, a)
    E4 = full4(tan[:, 0])
    assert np.allclose(sym6(np.einsum("ijkl,kl->ij", E4, e)), sig[:, 0], rtol=1e-14)                    # :19  σ ≈ ∂σ∂ε ⊡ ε
    _, tan180 = oracle.transversely_isotropic(p, eps, -a)
    assert np.array_equal(tan, tan180)                                                                   # :22-24 a → −a
    al = np.deg2rad(45)
    R = np.array([[1.0, 0, 0], [0, np.cos(al), np.sin(al)], [0, -np.sin(al), np.cos(al)]]).T             # :27-33  R_ij = e_i · b_j
    e45 = R.T @ e @ R
    _, tan45 = oracle.transversely_isotropic(p, sym6(0.5 * (e45 + e45.T)).reshape(6, 1), a)
    assert np.array_equal(tan45, tan)
    # :36-47 engineering compliance (Voigt order 11,22,33,23,13,12, engineering shear)
    E_L, E_T, G_LT, n_LT, n_TT = TI["E_L"], TI["E_T"], TI["G_LT"], TI["nu_LT"], TI["nu_TT"]
    G_TT = E_T / 2 / (1 + n_TT)
    Cref = np.array([[1 / E_L, -n_LT / E_L, -n_LT / E_L, 0, 0, 0], [-n_LT / E_L, 1 / E_T, -n_TT / E_T, 0, 0, 0], [-n_LT / E_L, -n_TT / E_T, 1 / E_T, 0, 0, 0],
                     [0, 0, 0, 1 / G_TT, 0, 0], [0, 0, 0, 0, 1 / G_LT, 0], [0, 0, 0, 0, 0, 1 / G_LT]])
    V = ((0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1))
    D = np.array([[E4[i, j, k, l] for (k, l) in V] for (i, j) in V])                                     # stiffness in Voigt
    assert np.allclose(np.linalg.inv(D), Cref, rtol=1e-12, atol=1e-20)


def test_twin_transviso_matches_oracle(oracle):
    tw = H.hosttwin()
    p = ti_params(oracle)
    N = 4000
    eps = sy.host_uniform(N, 6, 99, [-2e-3] * 6, [2e-3] * 6)
    a3 = unit_directions(N, 100)
    sig_ref, tan_ref = oracle.transversely_isotropic(p, eps, a3)
    sig, tan = np.zeros_like(sig_ref), np.zeros_like(tan_ref)
    assert tw.twin_transviso(C.byref(p), H.P(eps), H.P(a3), H.P(sig), H.P(tan), C.c_int64(N)) == 0
    assert H.rel_err(sig, sig_ref).max() < H.RTOL and H.rel_err(tan, tan_ref).max() < H.RTOL
    sig2 = np.zeros_like(sig)
    assert tw.twin_transviso(C.byref(p), H.P(eps), H.P(a3), H.P(sig2), None, C.c_int64(N)) == 0         # tangent optional
    assert np.array_equal(sig, sig2)


def test_energy_kats_and_trait_routes(oracle):
    Cd = np.array([np.e, 0, 0, 1.0, 0, 1.0]).reshape(6, 1)                                               # test_neohook.jl:5, test_yeoh.jl:6
    assert np.isclose(oracle.strain_energy_batch(abi.MM_NEOHOOK, (1.0, 1.0, 0, 0), Cd)[0], 0.5 * (np.e - 1) - 0.5 + 0.5 * 0.25, rtol=1e-14)
    ref = 0.5 * (np.e - 1) - (np.e - 1) ** 2 + (np.e - 1) ** 3 - 0.5 + 0.5 * 0.25
    assert np.isclose(oracle.strain_energy_batch(abi.MM_YEOH, (1.0, 1.0, -1.0, 1.0), Cd)[0], ref, rtol=1e-14)
    F = np.array([2.0, 0, 0, 1, 1, 0, 0, 0, 1]).reshape(9, 1)                                            # test_straintraits.jl:5, :51-54
    Fm = F[:, 0].reshape(3, 3, order="F")
    C6 = sym6(Fm.T @ Fm).reshape(6, 1)
    E6 = sym6((Fm.T @ Fm - np.eye(3)) / 2).reshape(6, 1)
    a = oracle.strain_energy_batch(abi.MM_NEOHOOK, (1.0, 1.0, 0, 0), C6, oracle.STRAIN_C)
    assert a == oracle.strain_energy_batch(abi.MM_NEOHOOK, (1.0, 1.0, 0, 0), E6, oracle.STRAIN_E) == oracle.strain_energy_batch(abi.MM_NEOHOOK, (1.0, 1.0, 0, 0), F, oracle.STRAIN_F)
    # LinearElastic: ψ = ½ ε:Eᵉ:ε = ½ σ:ε
    eps = sy.host_uniform(100, 6, 5, [-2e-3] * 6, [2e-3] * 6)
    psi = oracle.strain_energy_batch(-1, (200e3, 0.3), eps)
    sig = oracle.linear_elastic(200e3, 0.3, eps)[0]
    w = np.array([1, 2, 2, 1, 2, 1.0])[:, None]
    assert np.allclose(psi, 0.5 * (w * sig * eps).sum(0), rtol=1e-13)


@pytest.mark.parametrize("name,model,prm", HYPER)
def test_twin_energies_match_oracle(oracle, name, model, prm):
    tw = H.hosttwin()
    N = 5000
    nc, lo, hi, seed = sy.c3_defgrad()
    F = sy.host_uniform(N, nc, seed, lo, hi)
    Fm = F.reshape(3, 3, -1, order="F")
    Cm = np.einsum("kin,kjn->ijn", Fm, Fm)
    C6 = np.ascontiguousarray(np.stack([Cm[i, j] for i, j in SLOTS]))
    E6 = C6.copy()
    E6[[0, 3, 5]] -= 1.0
    E6 /= 2
    hp = abi.MMHyper(**prm)
    for strain, kind in ((C6, 0), (F, 1), (E6, 2)):
        ref = oracle.strain_energy_batch(model, hp, strain, kind)
        psi = np.zeros(N)
        assert tw.twin_hyper_energy(model, C.byref(hp), H.P(strain), kind, H.P(psi), C.c_int64(N)) == 0
        # ψ cancels (μ/2 (I₁−3) − μ ln J ≈ 0 near F = I): compare against the size of its terms
        scale = np.abs(ref) + abs(hp.mu) * np.abs(Cm[0, 0] + Cm[1, 1] + Cm[2, 2] - 3)
        assert (np.abs(psi - ref) / scale).max() < H.RTOL, (name, kind)
    eps = sy.host_uniform(N, 6, 5, [-2e-3] * 6, [2e-3] * 6)
    le = abi.MMLinearElastic(200e3, 0.3)
    psi = np.zeros(N)
    assert tw.twin_linear_elastic_energy(C.byref(le), H.P(eps), H.P(psi), C.c_int64(N)) == 0
    assert np.allclose(psi, oracle.strain_energy_batch(-1, (200e3, 0.3), eps), rtol=1e-13, atol=0)


def test_host_mirror_transviso_constructor_and_state(mm, oracle):
    m = mm.TransverselyIsotropic(E_L=TI["E_L"], E_T=TI["E_T"], G_LT=TI["G_LT"], ν_LT=TI["nu_LT"], ν_TT=TI["nu_TT"])
    p = ti_params(oracle)
    for f in ("L_perp", "L_par", "M_par", "G_perp", "G_par"):
        assert getattr(m._c, f) == getattr(p, f)
    assert mm.get_cache(m) is None                                                                        # test_transversely.jl:12-13
    st = mm.initial_material_state(m, (1.0, 0.0, 0.0), 5, device="cpu")
    assert st.N == 5 and st.data.shape == (3, 5) and np.array_equal(st.a3[:, 2], [1.0, 0.0, 0.0])
    st = mm.initial_material_state(m, 4, device="cpu")                                                    # default direction (1,0,0)
    assert st.N == 4 and np.array_equal(st.a3[:, 0], [1.0, 0.0, 0.0])
    with pytest.raises(ValueError, match="not a unit vector"):                                            # test_transversely.jl:49
        mm.initial_material_state(m, (-0.5, 0.0, 0.0), 3, device="cpu")
    with pytest.raises(RuntimeError, match="not the native strain measure"):
        mm.elastic_strain_energy_density(mm.LinearElastic(200e3, 0.3), mm.RightCauchyGreen(np.zeros((6, 1))))


@pytest.mark.gpu
@pytest.mark.parametrize("lname,lay", [("soa", 0), ("aos", 1)])
def test_gpu_transviso_matches_oracle(mm, cuda_lib, oracle, lname, lay):
    import torch

    N = 70001
    m = mm.TransverselyIsotropic(E_L=TI["E_L"], E_T=TI["E_T"], G_LT=TI["G_LT"], ν_LT=TI["nu_LT"], ν_TT=TI["nu_TT"])
    p = ti_params(oracle)
    eps = sy.host_uniform(N, 6, 99, [-2e-3] * 6, [2e-3] * 6)
    a3 = unit_directions(N, 100)
    sig_ref, tan_ref = oracle.transversely_isotropic(p, eps, a3)
    lay_of = (lambda a: a) if lay == 0 else (lambda a: np.ascontiguousarray(a.T))
    st = mm.initial_material_state(m, torch.from_numpy(lay_of(a3)).cuda(), layout=lname)
    sig, tan, st2 = mm.material_response(m, torch.from_numpy(lay_of(eps)).cuda(), st)
    assert st2 is st
    assert H.rel_err(H.soa(sig.cpu().numpy(), lay), sig_ref).max() < H.RTOL
    assert H.rel_err(H.soa(tan.cpu().numpy(), lay), tan_ref).max() < H.RTOL
    # reference's own identities through the CUDA path: a → −a leaves E unchanged; σ = E ⊡ ε
    st180 = mm.initial_material_state(m, torch.from_numpy(lay_of(-a3)).cuda(), layout=lname)
    _, tan180, _ = mm.material_response(m, torch.from_numpy(lay_of(eps)).cuda(), st180)
    assert torch.equal(tan, tan180)
    # host-buffer path (pipelined chunks) gives the same bits as the device path
    old = cuda_lib.mmh_set_chunk_points(16384)
    try:
        sth = mm.initial_material_state(m, lay_of(a3), device="cpu", layout=lname)
        sig_h, tan_h, _ = mm.material_response(m, lay_of(eps), sth)
    finally:
        cuda_lib.mmh_set_chunk_points(old)
    assert np.array_equal(sig_h, sig.cpu().numpy()) and np.array_equal(tan_h, tan.cpu().numpy())
    # one shared direction
    st1 = mm.initial_material_state(m, (0.0, 0.6, 0.8), N, layout=lname)
    s1, t1, _ = mm.material_response(m, torch.from_numpy(lay_of(eps)).cuda(), st1, tangent=False)
    a1 = np.repeat(np.array([[0.0], [0.6], [0.8]]), N, axis=1)
    assert t1 is None and H.rel_err(H.soa(s1.cpu().numpy(), lay), oracle.transversely_isotropic(p, eps, a1)[0]).max() < H.RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("lname,lay", [("soa", 0), ("aos", 1)])
def test_gpu_energies_match_oracle(mm, cuda_lib, oracle, lname, lay):
    import torch

    N = 50003
    lay_of = (lambda a: a) if lay == 0 else (lambda a: np.ascontiguousarray(a.T))
    nc, lo, hi, seed = sy.c3_defgrad()
    F = sy.host_uniform(N, nc, seed, lo, hi)
    Fm = F.reshape(3, 3, -1, order="F")
    Cm = np.einsum("kin,kjn->ijn", Fm, Fm)
    C6 = np.ascontiguousarray(np.stack([Cm[i, j] for i, j in SLOTS]))
    E6 = C6.copy()
    E6[[0, 3, 5]] -= 1.0
    E6 /= 2
    for name, model, prm in HYPER:
        m = getattr(mm, name)(λ=prm["lambda_"], μ=prm["mu"], c2=prm["c2"], c3=prm["c3"])
        hp = abi.MMHyper(**prm)
        scale = None
        for strain, kind, wrap in ((C6, 0, mm.RightCauchyGreen), (F, 1, mm.DeformationGradient), (E6, 2, mm.GreenLagrange), (C6, 0, None)):
            ref = oracle.strain_energy_batch(model, hp, strain, kind)
            x = torch.from_numpy(lay_of(strain)).cuda()
            psi = mm.elastic_strain_energy_density(m, wrap(x) if wrap else x, layout=lname).cpu().numpy()
            scale = np.abs(ref) + abs(hp.mu) * np.abs(Cm[0, 0] + Cm[1, 1] + Cm[2, 2] - 3)
            assert (np.abs(psi - ref) / scale).max() < H.RTOL, (name, kind)
        psi_h = mm.elastic_strain_energy_density(m, mm.DeformationGradient(lay_of(F)), layout=lname)      # host-buffer path
        psi_d = mm.elastic_strain_energy_density(m, mm.DeformationGradient(torch.from_numpy(lay_of(F)).cuda()), layout=lname)
        assert np.array_equal(psi_h, psi_d.cpu().numpy())
    eps = sy.host_uniform(N, 6, 5, [-2e-3] * 6, [2e-3] * 6)
    le = mm.LinearElastic(200e3, 0.3)
    psi = mm.elastic_strain_energy_density(le, torch.from_numpy(lay_of(eps)).cuda(), layout=lname).cpu().numpy()
    assert np.allclose(psi, oracle.strain_energy_batch(-1, (200e3, 0.3), eps), rtol=1e-13, atol=0)
    # ψ consistent with σ from material_response: ψ = ½ σ:ε
    sig, _, _ = mm.material_response(le, torch.from_numpy(lay_of(eps)).cuda(), layout=lname, tangent=False)
    w = np.array([1, 2, 2, 1, 2, 1.0])[:, None]
    assert np.allclose(psi, 0.5 * (w * H.soa(sig.cpu().numpy(), lay) * eps).sum(0), rtol=1e-12)
