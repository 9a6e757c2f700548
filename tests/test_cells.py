"""SURVEY §8(f) rank 4b — the caller's quadrature loop fused around material_response (mm_cells_*): for every cell
ε = sym(Σ_a u_a ⊗ ∇N_a), (σ, D, state) = material_response(m, ε, state), fe[a] += (σ ⋅ ∇N_a) dΩ.
The reference has no assembler (the loop is the caller's, e.g. Ferrite.jl's assemble_cell!), so the oracle's restatement of
that loop is pinned by finite-element identities: rigid-body motions give zero force, nodal forces of a cell sum to zero,
fe = Kₑ uₑ with Kₑ = Σ_q Bᵀ D B dΩ for LinearElastic, the constant-strain patch reproduces σ exactly, and per point it must
agree with the already-pinned point-wise oracle.  The CUDA path is then compared with that oracle."""
import numpy as np
import pytest

import helpers as H
import materialmodels_jl_b200.synth as sy

E, NU = 200e3, 0.3
AMP = {"tet4": 6e-4, "hex8": 9e-4}        # nodal displacement amplitudes giving a mixed elastic / plastic population
SLOTS = ((0, 0), (1, 0), (2, 0), (1, 1), (2, 1), (2, 2))


# ---- reference elements (numpy) -------------------------------------------------------------------
def tet4_geometry(X):
    """X (ncell, 4, 3) -> dNdx (ncell, 1, 4, 3), dΩ (ncell, 1): linear tetrahedron, 1-point rule"""
    dNdxi = np.array([[-1.0, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    J = np.einsum("cak,al->ckl", X, dNdxi)                        # J_kl = Σ_a X_ak ∂N_a/∂ξ_l
    detJ = np.linalg.det(J)
    dNdx = np.einsum("al,clk->cak", dNdxi, np.linalg.inv(J))
    return dNdx[:, None], (detJ / 6.0)[:, None]


def hex8_geometry(X):
    """X (ncell, 8, 3) -> dNdx (ncell, 8, 8, 3), dΩ (ncell, 8): trilinear hexahedron, 2x2x2 Gauss"""
    xi_n = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=float)
    g = 1.0 / np.sqrt(3.0)
    qp = xi_n * g
    dNdx, dO = [], []
    for q in range(8):
        d = np.stack([xi_n[:, k] * np.prod(1 + np.delete(xi_n, k, 1) * np.delete(qp[q], k)[None, :], axis=1) / 8 for k in range(3)], axis=1)   # (8,3)
        J = np.einsum("cak,al->ckl", X, d)
        dNdx.append(np.einsum("al,clk->cak", d, np.linalg.inv(J)))
        dO.append(np.linalg.det(J))
    return np.stack(dNdx, axis=1), np.stack(dO, axis=1)


def random_cells(kind, ncell, seed):
    rng = np.random.default_rng(seed)
    if kind == "tet4":
        X0 = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
        X = X0[None] + 0.15 * (2 * rng.random((ncell, 4, 3)) - 1)
        return X, 4, 1, tet4_geometry(X)
    X0 = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=float) * 0.5
    X = X0[None] + 0.1 * (2 * rng.random((ncell, 8, 3)) - 1)
    return X, 8, 8, hex8_geometry(X)


def pack(dNdx, dO, U):
    """(ncell,nq,nn,3), (ncell,nq), (ncell,nn,3) -> the library's component-major arrays"""
    ncell, nq, nn, _ = dNdx.shape
    d = np.ascontiguousarray(dNdx.reshape(ncell * nq, nn * 3).T)
    return d, np.ascontiguousarray(dO.reshape(ncell * nq)), np.ascontiguousarray(U.reshape(ncell, nn * 3).T)


def sym6(m):
    return np.array([m[i, j] for i, j in SLOTS])


@pytest.mark.parametrize("kind", ["tet4", "hex8"])
def test_oracle_cell_loop_finite_element_identities(oracle, kind):
    ncell = 40
    X, nn, nq, (dNdx, dO) = random_cells(kind, ncell, 5)
    assert (dO > 0).all()
    rng = np.random.default_rng(6)
    # (i) rigid-body motion u = w × x + t  ->  ε = 0, fe = 0
    w, t = rng.random(3), rng.random(3)
    U = np.cross(w[None, None, :], X) + t
    d, o, u = pack(dNdx, dO, U)
    r = oracle.cells(0, (E, NU), nn, nq, d, o, u)
    assert np.abs(r["sig"]).max() < 1e-9 * E and np.abs(r["fe"]).max() < 1e-9 * E
    # (ii) constant-strain patch u = G x: σ = Eᵉ:sym(G) at every point, and the cell's nodal forces sum to zero
    G = 1e-3 * (2 * rng.random((3, 3)) - 1)
    U = X @ G.T
    d, o, u = pack(dNdx, dO, U)
    r = oracle.cells(0, (E, NU), nn, nq, d, o, u)
    eps = sym6(0.5 * (G + G.T))
    s_ref, _ = oracle.linear_elastic(E, NU, eps.reshape(6, 1))
    assert np.allclose(r["sig"], s_ref, rtol=1e-11, atol=1e-9)
    fe = r["fe"].T.reshape(ncell, nn, 3)
    assert np.abs(fe.sum(axis=1)).max() < 1e-10 * np.abs(fe).max()
    # (iii) fe = Kₑ uₑ,  Kₑ = Σ_q Bᵀ D B dΩ  (LinearElastic), arbitrary nodal displacements
    U = 1e-3 * (2 * rng.random((ncell, nn, 3)) - 1)
    d, o, u = pack(dNdx, dO, U)
    r = oracle.cells(0, (E, NU), nn, nq, d, o, u)
    D36 = oracle.elastic_tangent_3d(E, NU)
    D = np.zeros((3, 3, 3, 3))
    for I, (i, j) in enumerate(SLOTS):
        for J, (k, l) in enumerate(SLOTS):
            D[i, j, k, l] = D[j, i, k, l] = D[i, j, l, k] = D[j, i, l, k] = D36[I + 6 * J]
    K = np.einsum("cqaj,ijkl,cqbl,cq->caibk", dNdx, D, dNdx, dO)
    assert np.allclose(fe_of(r, ncell, nn), np.einsum("caibk,cbk->cai", K, U), rtol=1e-11, atol=1e-12 * E)
    rk = oracle.cells(0, (E, NU), nn, nq, d, o, u, want_stiffness=True)                # the oracle's ke[i,j] += δε_i ⊡ D ⊡ Δε_j dΩ loop
    Ko = ke_of(rk, ncell, nn)
    assert np.allclose(Ko, K.reshape(ncell, 3 * nn, 3 * nn), rtol=1e-12, atol=1e-12 * E) and np.allclose(Ko, Ko.transpose(0, 2, 1), rtol=1e-12, atol=1e-12 * E)
    # (iv) per point the loop calls the (pinned) point-wise update: Plastic through the cell loop == Plastic on the same strains
    p = H.plastic_c()
    U = AMP[kind] * (2 * rng.random((ncell, nn, 3)) - 1)
    d, o, u = pack(dNdx, dO, U)
    P = ncell * nq
    rc = oracle.cells(1, p, nn, nq, d, o, u, np.zeros((14, P)))
    grad = np.einsum("cak,cqal->cqkl", U, dNdx).reshape(P, 3, 3)
    eps = np.stack([0.5 * (grad[:, i, j] + grad[:, j, i]) for i, j in SLOTS])
    rp = oracle.plastic(p, eps, np.zeros((14, P)))
    assert np.array_equal(rc["flags"], rp["flags"]) and np.array_equal(rc["iters"], rp["iters"]) and 0.1 < rc["flags"].mean() < 0.95
    assert H.rel_err(rc["sig"], rp["sig"]).max() < 1e-13 and H.rel_err(rc["tan"], rp["tan"]).max() < 1e-13 and np.allclose(rc["state"], rp["state"], rtol=1e-13, atol=1e-18)
    s = np.zeros((P, 3, 3))
    for I, (i, j) in enumerate(SLOTS):
        s[:, i, j] = s[:, j, i] = rc["sig"][I]
    f_ref = np.einsum("cqkl,cqal,cq->cak", s.reshape(ncell, nq, 3, 3), dNdx, dO)
    assert np.allclose(fe_of(rc, ncell, nn), f_ref, rtol=1e-12, atol=1e-12)
    # (v) consistent element stiffness with the (non-major-symmetric) J2 tangent of every point
    rk = oracle.cells(1, p, nn, nq, d, o, u, np.zeros((14, P)), want_stiffness=True)
    Dq = np.zeros((P, 3, 3, 3, 3))
    for I, (i, j) in enumerate(SLOTS):
        for J, (k, l) in enumerate(SLOTS):
            Dq[:, i, j, k, l] = Dq[:, j, i, k, l] = Dq[:, i, j, l, k] = Dq[:, j, i, l, k] = rc["tan"][I + 6 * J]
    Kp = np.einsum("cqaj,cqijkl,cqbl,cq->caibk", dNdx, Dq.reshape(ncell, nq, 3, 3, 3, 3), dNdx, dO).reshape(ncell, 3 * nn, 3 * nn)
    assert np.allclose(ke_of(rk, ncell, nn), Kp, rtol=1e-11, atol=1e-12 * E)


def fe_of(r, ncell, nn):
    return r["fe"].T.reshape(ncell, nn, 3)


def ke_of(r, ncell, nn):
    return r["ke"].T.reshape(ncell, 3 * nn, 3 * nn)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,ncell", [("tet4", 20011), ("hex8", 5003)])
def test_gpu_cell_loop_matches_oracle(mm, cuda_lib, oracle, kind, ncell):
    import torch

    X, nn, nq, (dNdx, dO) = random_cells(kind, ncell, 7)
    rng = np.random.default_rng(8)
    P = ncell * nq
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    p = H.plastic_c()
    st = mm.initial_material_state(m, P)
    st_ref = np.zeros((14, P))
    for step in range(2):                                             # second step starts from a non-virgin state
        U = AMP[kind] * (2 * rng.random((ncell, nn, 3)) - 1)
        d, o, u = pack(dNdx, dO, U)
        r = oracle.cells(1, p, nn, nq, d, o, u, st_ref)
        fe, new, ex = mm.cell_internal_forces(m, dev(d), dev(o), dev(u), st, tangent=True, stress=True)
        assert r["n_fail"] == 0 and 0.1 < r["flags"].mean() < 0.97
        assert np.array_equal(new.flags.cpu().numpy(), r["flags"]), "yield flags differ"
        assert np.array_equal(new.iters.cpu().numpy(), r["iters"]), "Newton counts differ"
        assert H.rel_err(ex["stress"].cpu().numpy(), r["sig"]).max() < H.RTOL and H.rel_err(ex["tangent"].cpu().numpy(), r["tan"]).max() < H.RTOL
        fscale = np.abs(r["fe"]).max()
        assert np.abs(fe.cpu().numpy() - r["fe"]).max() < H.RTOL * fscale
        assert np.abs(new.data.cpu().numpy() - r["state"]).max() < 1e-12 * max(1.0, np.abs(r["state"]).max())
        # element stiffness BᵀDB (D stays in registers)
        rk = oracle.cells(1, p, nn, nq, d, o, u, st_ref, want_stiffness=True)
        fe3, new3, ex3 = mm.cell_internal_forces(m, dev(d), dev(o), dev(u), st, stiffness=True)
        kscale = np.abs(rk["ke"]).max(axis=0)
        assert (np.abs(ex3["stiffness"].cpu().numpy() - rk["ke"]).max(axis=0) / kscale).max() < H.RTOL
        assert np.abs(fe3.cpu().numpy() - r["fe"]).max() < H.RTOL * fscale                      # (another kernel instantiation: not bitwise)
        assert np.abs(new3.data.cpu().numpy() - r["state"]).max() < 1e-12 * max(1.0, np.abs(r["state"]).max()) and torch.equal(new3.flags, new.flags)
        # the lean call (no σ, no tangent written) gives the same forces and state bit for bit
        fe2, new2, ex2 = mm.cell_internal_forces(m, dev(d), dev(o), dev(u), st)
        assert ex2["stress"] is None and ex2["tangent"] is None
        assert torch.equal(fe2, fe) and torch.equal(new2.data, new.data)
        st, st_ref = new, r["state"]
    le = mm.LinearElastic(E=E, ν=NU)
    r = oracle.cells(0, (E, NU), nn, nq, d, o, u)
    fe, _, ex = mm.cell_internal_forces(le, dev(d), dev(o), dev(u), stress=True)
    assert np.abs(fe.cpu().numpy() - r["fe"]).max() < H.RTOL * np.abs(r["fe"]).max() and H.rel_err(ex["stress"].cpu().numpy(), r["sig"]).max() < H.RTOL
    _, _, exk = mm.cell_internal_forces(le, dev(d), dev(o), dev(u), stiffness=True)
    Kg = exk["stiffness"].cpu().numpy().T.reshape(ncell, 3 * nn, 3 * nn)
    assert np.allclose(np.einsum("crs,cs->cr", Kg, u.T), fe.cpu().numpy().T, rtol=1e-10, atol=1e-10 * np.abs(r["fe"]).max())   # fe = Kₑ uₑ
    # equilibrium identity through the CUDA path: the nodal forces of every cell sum to zero
    f = fe.cpu().numpy().T.reshape(ncell, nn, 3)
    assert np.abs(f.sum(axis=1)).max() < 1e-10 * np.abs(f).max()
    with pytest.raises(RuntimeError, match="unsupported element"):
        mm.cell_internal_forces(le, dev(np.zeros((9, 6))), dev(np.ones(6)), dev(np.zeros((9, 2))))


# ---- total-Lagrangian cells: NeoHook / Yeoh / StVenant --------------------------------------------------------------------------
import materialmodels_jl_b200._abi as abi  # noqa: E402

HYPER = [("NeoHook", abi.MM_NEOHOOK, sy.NEOHOOK_PARAMS), ("Yeoh", abi.MM_YEOH, sy.YEOH_PARAMS), ("StVenant", abi.MM_STVENANT, sy.NEOHOOK_PARAMS)]


def _F_of(U, dNdx):
    ncell, nq = dNdx.shape[:2]
    return np.eye(3)[None, None] + np.einsum("cak,cqal->cqkl", U, dNdx)


@pytest.mark.parametrize("name,model,prm", HYPER)
@pytest.mark.parametrize("kind", ["tet4", "hex8"])
def test_oracle_hyper_cells_energy_consistency_and_objectivity(oracle, kind, name, model, prm):
    """fe = ∂W/∂u with W = Σ_q ψ(F_q) dΩ (the energy of src/FiniteStrain/*.jl:28-35), by central differences of the oracle's energy;
    a rigid rotation of the cell gives F = R, C = I, no stress, no force."""
    ncell = 6
    X, nn, nq, (dNdx, dO) = random_cells(kind, ncell, 15)
    rng = np.random.default_rng(16)
    hp = abi.MMHyper(**prm)
    U = 0.05 * (2 * rng.random((ncell, nn, 3)) - 1)
    d, o, u = pack(dNdx, dO, U)
    r = oracle.cells_hyper(model, hp, nn, nq, d, o, u)

    def W(Uv):
        F = _F_of(Uv, dNdx).reshape(ncell * nq, 3, 3)
        F9 = np.ascontiguousarray(F.transpose(2, 1, 0).reshape(9, -1))          # column-major per point: index i + 3j
        return (oracle.strain_energy_batch(model, hp, F9, oracle.STRAIN_F) * dO.reshape(-1)).reshape(ncell, nq).sum(axis=1)

    h = 1e-6
    fd = np.zeros((ncell, nn, 3))
    for a in range(nn):
        for k in range(3):
            Up, Um = U.copy(), U.copy()
            Up[:, a, k] += h
            Um[:, a, k] -= h
            fd[:, a, k] = (W(Up) - W(Um)) / (2 * h)
    assert np.allclose(fe_of(r, ncell, nn), fd, rtol=2e-7, atol=1e-8 * np.abs(fd).max())
    # objectivity: u = (R − I) x + t
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    U = X @ (R - np.eye(3)).T + np.array([0.3, -0.2, 0.1])
    d, o, u = pack(dNdx, dO, U)
    r = oracle.cells_hyper(model, hp, nn, nq, d, o, u)
    assert np.abs(r["Pt"]).max() < 1e-12 and np.abs(r["fe"]).max() < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name,model,prm", HYPER)
@pytest.mark.parametrize("kind,ncell", [("tet4", 20011), ("hex8", 5003)])
def test_gpu_hyper_cells_match_oracle(mm, cuda_lib, oracle, kind, ncell, name, model, prm):
    import torch

    X, nn, nq, (dNdx, dO) = random_cells(kind, ncell, 17)
    rng = np.random.default_rng(18)
    U = 0.05 * (2 * rng.random((ncell, nn, 3)) - 1)
    d, o, u = pack(dNdx, dO, U)
    hp = abi.MMHyper(**prm)
    r = oracle.cells_hyper(model, hp, nn, nq, d, o, u)
    m = getattr(mm, name)(λ=prm["lambda_"], μ=prm["mu"], c2=prm["c2"], c3=prm["c3"])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    fe, _, ex = mm.cell_internal_forces(m, dev(d), dev(o), dev(u), tangent=True, stress=True)
    assert H.rel_err(ex["stress"].cpu().numpy(), r["Pt"]).max() < H.RTOL and H.rel_err(ex["tangent"].cpu().numpy(), r["dPdF"]).max() < H.RTOL
    assert np.abs(fe.cpu().numpy() - r["fe"]).max() < H.RTOL * np.abs(r["fe"]).max()
    fe2, _, ex2 = mm.cell_internal_forces(m, dev(d), dev(o), dev(u))            # lean call: nothing but fe leaves the kernel
    assert ex2["stress"] is None and ex2["tangent"] is None and torch.equal(fe2, fe)


# ---- TransverselyIsotropic through the cell loop (per-point direction a3) --------------------------------------------------------
TI = dict(E_L=100e3, E_T=10e3, G_LT=5e3, nu_LT=0.4, nu_TT=0.3)


def _ti_dirs(P, seed):
    a = sy.host_uniform(P, 3, seed, [-1.0] * 3, [1.0] * 3)
    a[:, np.linalg.norm(a, axis=0) < 1e-3] = np.array([[1.0], [0.0], [0.0]])
    return np.ascontiguousarray(a / np.linalg.norm(a, axis=0))


@pytest.mark.parametrize("kind", ["tet4", "hex8"])
def test_oracle_transviso_cells(oracle, kind):
    """per point equal to the point-wise TransverselyIsotropic oracle; fe = Kₑ uₑ with the oracle's own element stiffness"""
    ncell = 30
    X, nn, nq, (dNdx, dO) = random_cells(kind, ncell, 25)
    P = ncell * nq
    p = oracle.transviso_params(TI["E_L"], TI["E_T"], TI["G_LT"], TI["nu_LT"], TI["nu_TT"])
    a3 = _ti_dirs(P, 26)
    U = 1e-3 * (2 * np.random.default_rng(27).random((ncell, nn, 3)) - 1)
    d, o, u = pack(dNdx, dO, U)
    r = oracle.cells(2, p, nn, nq, d, o, u, a3, want_stiffness=True)
    grad = np.einsum("cak,cqal->cqkl", U, dNdx).reshape(P, 3, 3)
    eps = np.stack([0.5 * (grad[:, i, j] + grad[:, j, i]) for i, j in SLOTS])
    s_ref, t_ref = oracle.transversely_isotropic(p, eps, a3)
    assert H.rel_err(r["sig"], s_ref).max() < 1e-13 and H.rel_err(r["tan"], t_ref).max() < 1e-13
    K = ke_of(r, ncell, nn)
    assert np.allclose(np.einsum("crs,cs->cr", K, u.T), r["fe"].T, rtol=1e-11, atol=1e-12 * np.abs(r["fe"]).max())
    assert np.allclose(K, K.transpose(0, 2, 1), rtol=1e-11, atol=1e-9)             # hyperelastic stiffness: major-symmetric


@pytest.mark.gpu
@pytest.mark.parametrize("kind,ncell", [("tet4", 20011), ("hex8", 5003)])
def test_gpu_transviso_cells_match_oracle(mm, cuda_lib, oracle, kind, ncell):
    import torch

    X, nn, nq, (dNdx, dO) = random_cells(kind, ncell, 28)
    P = ncell * nq
    p = oracle.transviso_params(TI["E_L"], TI["E_T"], TI["G_LT"], TI["nu_LT"], TI["nu_TT"])
    m = mm.TransverselyIsotropic(E_L=TI["E_L"], E_T=TI["E_T"], G_LT=TI["G_LT"], ν_LT=TI["nu_LT"], ν_TT=TI["nu_TT"])
    a3 = _ti_dirs(P, 29)
    U = 1e-3 * (2 * np.random.default_rng(30).random((ncell, nn, 3)) - 1)
    d, o, u = pack(dNdx, dO, U)
    r = oracle.cells(2, p, nn, nq, d, o, u, a3, want_stiffness=True)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    st = mm.initial_material_state(m, dev(a3))
    fe, st2, ex = mm.cell_internal_forces(m, dev(d), dev(o), dev(u), st, tangent=True, stress=True, stiffness=True)
    assert st2 is st
    assert H.rel_err(ex["stress"].cpu().numpy(), r["sig"]).max() < H.RTOL and H.rel_err(ex["tangent"].cpu().numpy(), r["tan"]).max() < H.RTOL
    assert np.abs(fe.cpu().numpy() - r["fe"]).max() < H.RTOL * np.abs(r["fe"]).max()
    kscale = np.abs(r["ke"]).max(axis=0)
    assert (np.abs(ex["stiffness"].cpu().numpy() - r["ke"]).max(axis=0) / kscale).max() < H.RTOL
