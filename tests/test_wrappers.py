"""Dimension wrappers (reference src/wrappers.jl): the oracle against the reference's own tests
(test/test_wrappers.jl), the kernels' math on the host against the oracle, and the CUDA path against the oracle."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
import materialmodels_jl_b200._abi as abi
import materialmodels_jl_b200.synth as sy

E, NU = 200e3, 0.3
KINDS = [1, 2, 3, 4, 5]          # UniaxialStrain, PlaneStrain, UniaxialStress, PlaneStress, ShellStress
NC = {1: 1, 2: 3, 3: 1, 4: 3, 5: 6}
# strain amplitudes giving a mixed elastic / plastic population per kind
AMPL = {1: 2.5e-3, 2: 1.0e-3, 3: 2.5e-3, 4: 1.2e-3, 5: 0.8e-3}
# UniaxialStress: the 1x1 tangent is a Schur complement through a 5x5 block (2.7e5 -> O(1e2..1e5)): cond·eps ~ 1e-11
TAN_TOL = {1: H.RTOL, 2: H.RTOL, 3: 1e-10, 4: H.RTOL, 5: H.RTOL}


def sig_err(a, b):
    """stress error against the problem's stress scale: a reduced stress can have a single component that passes through
    zero (UniaxialStrain / UniaxialStress), where an error relative to its own magnitude is meaningless"""
    scale = np.maximum(np.linalg.norm(b, axis=0), sy.PLASTIC_PARAMS["sigma_y"])
    return (np.linalg.norm(a - b, axis=0) / scale).max()


# ---- oracle vs the reference's KATs (test_wrappers.jl) ---------------------------------------------
def test_oracle_wrappers_reference_kats(oracle):
    rng = np.random.default_rng(11)
    e = rng.random((1, 5))                                                           # :10-14 UniaxialStress
    r = oracle.wrapped_linear_elastic(oracle.UNIAXIAL_STRESS, E, NU, e)
    assert np.allclose(r["tan"][0], E) and np.allclose(r["sig"], E * e) and r["n_fail"] == 0
    e = rng.random((3, 5))                                                           # :16-21 PlaneStress
    r = oracle.wrapped_linear_elastic(oracle.PLANE_STRESS, E, NU, e)
    vo = [0, 2, 1]                                                                   # storage (11,21,22) -> Voigt (11,22,12)
    ref = E / (1 - NU**2) * np.array([[1, NU, 0], [NU, 1, 0], [0, 0, (1 - NU) / 2]])
    W = np.array([1.0, 2.0, 1.0])
    for j in range(5):
        T = r["tan"][:, j].reshape(3, 3, order="F")
        assert np.allclose(T[np.ix_(vo, vo)], ref, rtol=1e-12, atol=1e-9)
        assert np.allclose(r["sig"][:, j], T @ (W * e[:, j]), rtol=1e-12)             # σ ≈ ∂σ∂ε ⊡ Δε
    e = rng.random((6, 5))                                                           # :23-38 ShellStress
    r = oracle.wrapped_linear_elastic(oracle.SHELL_STRESS, E, NU, e)
    a = (1 - NU) / 2
    ref = E / (1 - NU**2) * np.array([[1, NU, 0, 0, 0, 0], [NU, 1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0], [0, 0, 0, a, 0, 0], [0, 0, 0, 0, a, 0], [0, 0, 0, 0, 0, a]])
    vo6 = [0, 3, 5, 4, 2, 1]
    W6 = np.array([1.0, 2, 2, 1, 2, 1])
    for j in range(5):
        T = r["tan"][:, j].reshape(6, 6, order="F")
        assert np.allclose(T[np.ix_(vo6, vo6)], ref, rtol=1e-12, atol=1e-9)
        assert abs(r["sig"][5, j]) < 1e-10 * E
        assert np.allclose(r["sig"][:, j], T @ (W6 * e[:, j]), rtol=1e-11, atol=1e-9)
    e = rng.random((1, 5))                                                           # :44-50 UniaxialStrain
    r = oracle.wrapped_linear_elastic(oracle.UNIAXIAL_STRAIN, E, NU, e)
    e3 = np.zeros((6, 5))
    e3[0] = e[0]
    s3, t3 = oracle.linear_elastic(E, NU, e3)
    assert np.array_equal(r["sig"][0], s3[0]) and np.array_equal(r["tan"][0], t3[0])
    e = rng.random((3, 5))                                                           # :52-56 PlaneStrain
    r = oracle.wrapped_linear_elastic(oracle.PLANE_STRAIN, E, NU, e)
    ref = E / (1 + NU) / (1 - 2 * NU) * np.array([[1 - NU, NU, 0], [NU, 1 - NU, 0], [0, 0, (1 - 2 * NU) / 2]])
    for j in range(5):
        T = r["tan"][:, j].reshape(3, 3, order="F")
        assert np.allclose(T[np.ix_(vo, vo)], ref, rtol=1e-12, atol=1e-9)
        assert np.allclose(r["sig"][:, j], T @ (W * e[:, j]), rtol=1e-12)


def test_oracle_plane_stress_plastic_properties(oracle):
    """Plastic under PlaneStress / UniaxialStress / ShellStress: the converged 3D stress has vanishing out-of-plane
    components, the state is the inner model's, and the tangent is consistent with a finite difference of the stress."""
    p = H.plastic_c()
    N = 2000
    for kind in (3, 4, 5):
        nc = NC[kind]
        eps = sy.host_uniform(N, nc, 900 + kind, -AMPL[kind], AMPL[kind])
        r = oracle.wrapped_plastic(kind, p, eps, np.zeros((14, N)))
        assert r["n_fail"] == 0 and 0.2 < (r["flags"] & 1).mean() < 0.9
        h = 1e-9
        for c in range(nc):
            ep = eps.copy()
            ep[c] += h
            rp = oracle.wrapped_plastic(kind, p, ep, np.zeros((14, N)))
            same = rp["flags"] == r["flags"]
            fd = (rp["sig"] - r["sig"]) / h
            col = r["tan"][c * nc:(c + 1) * nc] * (1.0 if c in {1: (0,), 3: (0, 2), 6: (0, 3, 5)}[nc] else 2.0)   # off-diagonal strain slots count twice
            err = np.linalg.norm(fd[:, same] - col[:, same], axis=0) / np.linalg.norm(col[:, same], axis=0).clip(1.0)
            assert np.percentile(err, 99) < 2e-3      # FD noise: inner ftol and plane-stress tol are 1e-8 on stresses, h = 1e-9


# ---- kernels' math on the host vs the oracle -------------------------------------------------------
@pytest.mark.parametrize("kind", KINDS)
def test_twin_wrapped_matches_oracle(oracle, kind):
    tw = H.hosttwin()
    N = 6000
    nc = NC[kind]
    p = H.plastic_c()
    st = np.zeros((14, N))
    for step in range(2):
        eps = sy.host_uniform(N, nc, 777 + kind + 10 * step, -AMPL[kind], AMPL[kind])
        r = oracle.wrapped_plastic(kind, p, eps, st)
        sig, tan, so = np.zeros((nc, N)), np.zeros((nc * nc, N)), np.zeros((14, N))
        fl, it, oi = np.zeros(N, np.uint8), np.zeros(N, np.uint16), np.zeros(N, np.uint8)
        tw.twin_wrapped_plastic(kind, C.byref(p), H.P(eps), H.P(st), H.P(sig), H.P(tan), H.P(so), H.P(fl), H.P(it), H.P(oi),
                                C.c_double(1e-8), 1000, C.c_double(1e-8), 10, C.c_int64(N))
        assert r["n_fail"] == 0 and 0.15 < (r["flags"] & 1).mean() < 0.9
        assert np.array_equal(fl, r["flags"]) and np.array_equal(it, r["iters"]) and np.array_equal(oi, r["outer_iters"])
        assert sig_err(sig, r["sig"]) < H.RTOL and H.rel_err(tan, r["tan"]).max() < TAN_TOL[kind]
        assert np.abs(so - r["state"]).max() < 1e-12 * max(1.0, np.abs(r["state"]).max())
        if kind >= 3:        # the state machine the lane-persistent kernel drives (mm_wrap_lp.cuh): same counts, same flags, same values
            sig2, tan2, so2 = np.zeros((nc, N)), np.zeros((nc * nc, N)), np.zeros((14, N))
            fl2, it2, oi2 = np.zeros(N, np.uint8), np.zeros(N, np.uint16), np.zeros(N, np.uint8)
            tw.twin_wrapped_plastic_lp(kind, C.byref(p), H.P(eps), H.P(st), H.P(sig2), H.P(tan2), H.P(so2), H.P(fl2), H.P(it2), H.P(oi2),
                                       C.c_double(1e-8), 1000, C.c_double(1e-8), 10, C.c_int64(N))
            assert np.array_equal(fl2, r["flags"]) and np.array_equal(it2, r["iters"]) and np.array_equal(oi2, r["outer_iters"])
            assert sig_err(sig2, r["sig"]) < H.RTOL and H.rel_err(tan2, r["tan"]).max() < TAN_TOL[kind]
            assert np.abs(so2 - r["state"]).max() < 1e-12 * max(1.0, np.abs(r["state"]).max())
            assert np.abs(sig2 - sig).max() <= 1e-13 * np.abs(sig).max() and np.abs(so2 - so).max() <= 1e-13 * max(1.0, np.abs(so).max())
        st = r["state"]
    e = sy.host_uniform(500, nc, 5, -1e-3, 1e-3)
    rr = oracle.wrapped_linear_elastic(kind, E, NU, e)
    s2, t2, o2 = np.zeros((nc, 500)), np.zeros((nc * nc, 500)), np.zeros(500, np.uint8)
    tw.twin_wrapped_elastic(kind, C.byref(abi.MMLinearElastic(E, NU)), H.P(e), H.P(s2), H.P(t2), H.P(o2), C.c_int64(500))
    assert H.rel_err(s2, rr["sig"]).max() < H.RTOL and H.rel_err(t2, rr["tan"]).max() < H.RTOL and np.array_equal(o2, rr["outer_iters"])


# ---- CUDA path ------------------------------------------------------------------------------------
DIMS = {1: "UniaxialStrain", 2: "PlaneStrain", 3: "UniaxialStress", 4: "PlaneStress", 5: "ShellStress"}


@pytest.mark.gpu
@pytest.mark.parametrize("lname,lay", [("soa", 0), ("aos", 1)])
@pytest.mark.parametrize("kind", KINDS)
def test_gpu_wrapped_plastic_and_elastic(mm, cuda_lib, oracle, kind, lname, lay):
    import torch

    N = 50021
    nc = NC[kind]
    dim = getattr(mm, DIMS[kind])()
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    p = H.plastic_c()
    st = mm.initial_material_state(m, N, layout=lname)
    st_ref = np.zeros((14, N) if lay == 0 else (N, 14))
    soa = (lambda a: a) if lay == 0 else (lambda a: np.ascontiguousarray(a.T))
    for step in range(2):
        eps = sy.host_uniform(N, nc, 777 + kind + 10 * step, -AMPL[kind], AMPL[kind], layout=lay)
        r = oracle.wrapped_plastic(kind, p, eps, st_ref, layout=lay)
        s, t, new = mm.material_response(dim, m, torch.from_numpy(eps).cuda(), st, cache=mm.get_cache(m))
        assert r["n_fail"] == 0
        assert np.array_equal(new.flags.cpu().numpy(), r["flags"]), "yield flags differ"
        assert np.array_equal(new.iters.cpu().numpy(), r["iters"]), "inner Newton counts differ"
        assert np.array_equal(new.outer_iters.cpu().numpy(), r["outer_iters"]), "outer (plane-stress) iteration counts differ"
        assert sig_err(soa(s.cpu().numpy()), soa(r["sig"])) < H.RTOL
        assert H.rel_err(soa(t.cpu().numpy()), soa(r["tan"])).max() < TAN_TOL[kind]
        assert np.abs(new.data.cpu().numpy() - r["state"]).max() < 1e-12 * max(1.0, np.abs(r["state"]).max())
        st, st_ref = new, r["state"]
    le = mm.LinearElastic(E=E, ν=NU)
    e = sy.host_uniform(N, nc, 5, -1e-3, 1e-3, layout=lay)
    rr = oracle.wrapped_linear_elastic(kind, E, NU, e, layout=lay)
    s, t, _ = mm.material_response(dim, le, torch.from_numpy(e).cuda(), None, layout=lname)
    assert H.rel_err(soa(s.cpu().numpy()), soa(rr["sig"])).max() < H.RTOL and H.rel_err(soa(t.cpu().numpy()), soa(rr["tan"])).max() < H.RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("lname,lay", [("soa", 0), ("aos", 1)])
@pytest.mark.parametrize("kind", [3, 4, 5])
def test_gpu_wrapped_plastic_lane_persistent_equals_thread_per_point(mm, cuda_lib, kind, lname, lay):
    """The lane-persistent kernel (default for the stress-restricted kinds) against the one-thread-per-point kernel
    (mm_set_tuning("wrapped_j2_variant", 0)): flags, inner and outer counts identical, values to rounding — whatever the section
    thresholds and the chunk length, for ragged sizes, without a tangent, and updating the state in place."""
    import torch

    nc = NC[kind]
    dim = getattr(mm, DIMS[kind])()
    m = mm.Plastic(**sy.PLASTIC_PARAMS)

    def run(N, variant, knobs=(), tangent=True, inplace=False):
        olds = [(b"wrapped_j2_variant", cuda_lib.mm_set_tuning(b"wrapped_j2_variant", variant))]
        for name, v in knobs:
            olds.append((name, cuda_lib.mm_set_tuning(name, v)))
        try:
            st = mm.initial_material_state(m, N, layout=lname)
            outs = []
            for step in range(2):
                eps = torch.from_numpy(sy.host_uniform(N, nc, 900 + kind + 10 * step, -AMPL[kind], AMPL[kind], layout=lay)).cuda()
                if inplace:        # state_out == state_in through the C ABI (include/mmb200.h allows it)
                    P = lambda t_: C.c_void_p(t_.data_ptr())
                    shp = (lambda n_: (n_, N)) if lay == 0 else (lambda n_: (N, n_))
                    s = torch.empty(shp(nc), dtype=torch.float64, device="cuda")
                    t = torch.empty(shp(nc * nc), dtype=torch.float64, device="cuda")
                    fl, it_ = torch.empty(N, dtype=torch.uint8, device="cuda"), torch.empty(N, dtype=torch.int16, device="cuda")
                    oi, nf = torch.empty(N, dtype=torch.uint8, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
                    data = st.data.clone()
                    rc = cuda_lib.mm_wrapped_plastic(kind, C.byref(m._c), P(eps), P(data), P(s), P(t), P(data), P(fl), P(it_), P(oi), P(nf), None, None, N, lay, None)
                    assert rc == 0
                    torch.cuda.synchronize()
                    new = mm.PlasticState(data, lname)
                    outs.append((s.cpu().numpy(), t.cpu().numpy(), data.cpu().numpy().copy(), fl.cpu().numpy(), it_.cpu().numpy().view(np.uint16), oi.cpu().numpy()))
                    st = new
                    continue
                s, t, new = mm.material_response(dim, m, eps, st, tangent=tangent)
                outs.append((s.cpu().numpy(), None if t is None else t.cpu().numpy(), new.data.cpu().numpy().copy(), new.flags.cpu().numpy(), new.iters.cpu().numpy(),
                             new.outer_iters.cpu().numpy()))
                st = new
            return outs
        finally:
            for name, v in olds:
                cuda_lib.mm_set_tuning(name, v)

    def same(a, b):
        for (s1, t1, d1, f1, i1, o1), (s2, t2, d2, f2, i2, o2) in zip(a, b):
            assert np.array_equal(f1, f2) and np.array_equal(i1, i2) and np.array_equal(o1, o2)
            assert np.abs(s1 - s2).max() <= 1e-13 * np.abs(s1).max() and np.abs(d1 - d2).max() <= 1e-13 * max(1.0, np.abs(d1).max())
            if t1 is not None:
                assert np.abs(t1 - t2).max() <= 1e-12 * np.abs(t1).max()

    for N in (1, 31, 50021):
        ref = run(N, 0)
        assert (ref[1][3] & 1).mean() > 0.1 or N < 100
        same(ref, run(N, 1))
    N = 20011
    ref = run(N, 0)
    for knobs in (((b"wrapped_j2_th_outer", 1), (b"wrapped_j2_th_final", 1)), ((b"wrapped_j2_th_outer", 32), (b"wrapped_j2_th_final", 32)),
                  ((b"wrapped_j2_chunk", 32),), ((b"wrapped_j2_chunk", 4096), (b"wrapped_j2_th_final", 3))):
        same(ref, run(N, 1, knobs))
    same(run(N, 0, tangent=False), run(N, 1, tangent=False))
    same(ref, run(N, 1, inplace=True))


@pytest.mark.gpu
def test_gpu_plane_stress_failure_raises_reference_error(mm, cuda_lib):
    import torch

    N = 4000
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    eps = torch.from_numpy(sy.host_uniform(N, 3, 42, -2e-3, 2e-3)).cuda()
    with pytest.raises(RuntimeError, match="Could not find plane stress state"):       # wrappers.jl:78
        mm.material_response(mm.PlaneStress(), m, eps, mm.initial_material_state(m, N), options={"plane_stress_max_iter": 1})
    with pytest.raises(RuntimeError, match="Material model not converged"):             # Plastic.jl:157 (inner)
        mm.material_response(mm.PlaneStress(), m, eps, mm.initial_material_state(m, N), options={"nlsolve_params": {"iterations": 0}})


# ---- wrappers around TransverselyIsotropic (wrappers.jl is generic in the material) and host-buffer forms --------------
TI = dict(E_L=100e3, E_T=10e3, G_LT=5e3, nu_LT=0.4, nu_TT=0.3)          # test_transversely.jl:4-8


def _unit_dirs(N, seed):
    a = sy.host_uniform(N, 3, seed, [-1.0] * 3, [1.0] * 3)
    a[:, np.linalg.norm(a, axis=0) < 1e-3] = np.array([[1.0], [0.0], [0.0]])
    return np.ascontiguousarray(a / np.linalg.norm(a, axis=0))


def test_oracle_wrapped_transviso_is_the_schur_complement(oracle):
    """PlaneStress around a linear model: one outer update, tangent = Schur complement of the 3D stiffness (wrappers.jl:63-70),
    and the converged 3D strain gives vanishing out-of-plane stresses."""
    p = oracle.transviso_params(TI["E_L"], TI["E_T"], TI["G_LT"], TI["nu_LT"], TI["nu_TT"])
    N = 50
    a3 = _unit_dirs(N, 3)
    e = sy.host_uniform(N, 3, 4, [-1e-3] * 3, [1e-3] * 3)
    r = oracle.wrapped_transviso(oracle.PLANE_STRESS, p, e, a3)
    assert r["n_fail"] == 0 and np.all(r["outer_iters"] == 1)
    _, E3 = oracle.transversely_isotropic(p, np.zeros((6, N)), a3)
    slot = {(0, 0): 0, (1, 0): 1, (2, 0): 2, (1, 1): 3, (2, 1): 4, (2, 2): 5}
    inp, oop = [0, 1, 3], [2, 4, 5]                       # storage slots (11,21,22) in plane; (31,32,33) out of plane
    w = np.array([1.0, 2.0, 2.0, 1.0, 2.0, 1.0])         # ⊡ weights of the symmetric storage
    for j in range(N):
        T = E3[:, j].reshape(6, 6, order="F") * w[None, :]                          # σ_I = Σ_J T[I,J] ε_J on storage components
        S = T[np.ix_(inp, inp)] - T[np.ix_(inp, oop)] @ np.linalg.solve(T[np.ix_(oop, oop)], T[np.ix_(oop, inp)])
        assert np.allclose(r["sig"][:, j], S @ e[:, j], rtol=1e-11, atol=1e-12)
        Tred = r["tan"][:, j].reshape(3, 3, order="F") * np.array([1.0, 2.0, 1.0])[None, :]
        assert np.allclose(Tred, S, rtol=1e-11, atol=1e-7)


@pytest.mark.parametrize("kind", KINDS)
def test_twin_wrapped_transviso_matches_oracle(oracle, kind):
    tw = H.hosttwin()
    p = oracle.transviso_params(TI["E_L"], TI["E_T"], TI["G_LT"], TI["nu_LT"], TI["nu_TT"])
    N, nc = 3000, NC[kind]
    a3 = _unit_dirs(N, 30 + kind)
    e = sy.host_uniform(N, nc, 40 + kind, [-1e-3] * nc, [1e-3] * nc)
    r = oracle.wrapped_transviso(kind, p, e, a3)
    sig, tan, oit = np.zeros((nc, N)), np.zeros((nc * nc, N)), np.zeros(N, np.uint8)
    assert tw.twin_wrapped_transviso(kind, C.byref(p), H.P(e), H.P(a3), H.P(sig), H.P(tan), H.P(oit), C.c_int64(N)) == 0
    assert np.array_equal(oit, r["outer_iters"])
    scale = np.maximum(np.linalg.norm(r["sig"], axis=0), 1e-3 * TI["E_T"])
    assert (np.linalg.norm(sig - r["sig"], axis=0) / scale).max() < 1e-11
    assert H.rel_err(tan, r["tan"]).max() < 1e-10        # Schur complements through up to a 5x5 block of a stiffness with E_L/E_T = 10


@pytest.mark.gpu
@pytest.mark.parametrize("lname,lay", [("soa", 0), ("aos", 1)])
@pytest.mark.parametrize("kind", KINDS)
def test_gpu_wrapped_transviso_and_host_buffer_forms(mm, cuda_lib, oracle, kind, lname, lay):
    import torch

    N, nc = 30011, NC[kind]
    dim = getattr(mm, DIMS[kind])()
    m = mm.TransverselyIsotropic(E_L=TI["E_L"], E_T=TI["E_T"], G_LT=TI["G_LT"], ν_LT=TI["nu_LT"], ν_TT=TI["nu_TT"])
    p = oracle.transviso_params(TI["E_L"], TI["E_T"], TI["G_LT"], TI["nu_LT"], TI["nu_TT"])
    soa = (lambda a: a) if lay == 0 else (lambda a: np.ascontiguousarray(a.T))
    lay_of = soa
    a3 = _unit_dirs(N, 30 + kind)
    e = sy.host_uniform(N, nc, 40 + kind, [-1e-3] * nc, [1e-3] * nc)
    r = oracle.wrapped_transviso(kind, p, e, a3)
    st = mm.initial_material_state(m, torch.from_numpy(lay_of(a3)).cuda(), layout=lname)
    s, t, new = mm.material_response(dim, m, torch.from_numpy(lay_of(e)).cuda(), st)
    assert np.array_equal(new.outer_iters.cpu().numpy(), r["outer_iters"])
    scale = np.maximum(np.linalg.norm(r["sig"], axis=0), 1e-3 * TI["E_T"])
    assert (np.linalg.norm(soa(s.cpu().numpy()) - r["sig"], axis=0) / scale).max() < 1e-11
    assert H.rel_err(soa(t.cpu().numpy()), r["tan"]).max() < 1e-10
    # host-buffer forms (pipelined chunks, ragged last chunk) == device forms, bit for bit
    old = cuda_lib.mmh_set_chunk_points(7000)
    try:
        sth = mm.initial_material_state(m, lay_of(a3), device="cpu", layout=lname)
        sh, th, _ = mm.material_response(dim, m, lay_of(e), sth)
        assert np.array_equal(sh, s.cpu().numpy()) and np.array_equal(th, t.cpu().numpy())
        pm = mm.Plastic(**sy.PLASTIC_PARAMS)
        ep = sy.host_uniform(N, nc, 777 + kind, -AMPL[kind], AMPL[kind], layout=lay)
        sd, td, nd = mm.material_response(dim, pm, torch.from_numpy(ep).cuda(), mm.initial_material_state(pm, N, layout=lname))
        sh, th, nh = mm.material_response(dim, pm, ep, mm.initial_material_state(pm, N, device="cpu", layout=lname))
        assert np.array_equal(sh, sd.cpu().numpy()) and np.array_equal(th, td.cpu().numpy()) and np.array_equal(nh.data, nd.data.cpu().numpy())
        assert np.array_equal(nh.flags, nd.flags.cpu().numpy()) and np.array_equal(nh.outer_iters, nd.outer_iters.cpu().numpy())
        le = mm.LinearElastic(E=E, ν=NU)
        sd, td, _ = mm.material_response(dim, le, torch.from_numpy(ep).cuda(), None, layout=lname)
        sh, th, _ = mm.material_response(dim, le, ep, None, layout=lname)
        assert np.array_equal(sh, sd.cpu().numpy()) and np.array_equal(th, td.cpu().numpy())
    finally:
        cuda_lib.mmh_set_chunk_points(old)


# ---- wrappers around CrystalViscoPlastic (the last material of the generic wrappers.jl) ---------------------------------
XTAL_AMPL = {1: 6e-3, 2: 4e-3, 3: 6e-3, 4: 4e-3, 5: 3e-3}


def _xtal_compare(r, active, iters, oit, sig, tan, state, kind):
    """wrapped crystal vs oracle: bit-exact masks / iteration counts and 1e-12 values away from the reference Newton's chaotic
    trajectories (helpers.crystal_compare explains those); the outer loop multiplies inner solves, so `iters` (the LAST inner
    solve's count) >= CHAOTIC_ITERS or any failure marks the point as chaotic."""
    active = np.asarray(active).view(np.uint16)
    chaotic = ((r["active"] & 0xA000) != 0) | (r["iters"] >= H.CHAOTIC_ITERS) | ((active & 0xA000) != 0) | (np.asarray(iters) >= H.CHAOTIC_ITERS)
    ok = ~chaotic
    assert chaotic.mean() < 0.03, chaotic.mean()
    assert np.array_equal(active[ok], r["active"][ok])
    assert np.array_equal(np.asarray(iters)[ok], r["iters"][ok]) and np.array_equal(np.asarray(oit)[ok], r["outer_iters"][ok])
    assert sig_err(sig[:, ok], r["sig"][:, ok]) < H.RTOL
    assert H.rel_err(tan[:, ok], r["tan"][:, ok]).max() < max(TAN_TOL[kind], 1e-10)
    assert H.rel_err(state[:, ok], r["state"][:, ok]).max() < H.RTOL
    return int(chaotic.sum())


def test_oracle_wrapped_crystal_properties(oracle):
    """PlaneStress / UniaxialStress around the crystal: out-of-plane stresses of the stored 3D state vanish (wrappers.jl:57-61), and a
    UniaxialStrain step equals the 3D model at ε = (ε11, 0, …) (wrappers.jl:29-35)."""
    p = H.crystal_c(oracle)
    N = 300
    for kind, oop in ((4, [2, 4, 5]), (3, [1, 2, 3, 4, 5]), (5, [5])):
        e = sy.host_uniform(N, NC[kind], 50 + kind, -XTAL_AMPL[kind], XTAL_AMPL[kind])
        r = oracle.wrapped_crystal(kind, p, e, np.zeros((42, N)))
        good = (r["active"] & 0xA000) == 0
        assert good.mean() > 0.97 and (r["active"][good] & 0xFFF).astype(bool).mean() > 0.2
        assert np.abs(r["state"][oop][:, good]).max() < 1e-8 * 50      # tol · a stress scale
    e = sy.host_uniform(N, 1, 60, -XTAL_AMPL[1], XTAL_AMPL[1])
    r = oracle.wrapped_crystal(1, p, e, np.zeros((42, N)))
    e3 = np.zeros((6, N)); e3[0] = e[0]
    r3 = oracle.crystal(p, e3, np.zeros((42, N)))
    assert np.array_equal(r["sig"][0], r3["sig"][0]) and np.array_equal(r["state"], r3["state"]) and np.array_equal(r["tan"][0], r3["tan"][0])
    assert np.array_equal(r["active"], r3["active"]) and np.array_equal(r["iters"], r3["iters"])


@pytest.mark.parametrize("kind", KINDS)
def test_twin_wrapped_crystal_matches_oracle(oracle, kind):
    tw = H.hosttwin()
    p = H.crystal_c(oracle)
    N = 400
    nc = NC[kind]
    st = np.zeros((42, N))
    nch = 0
    for step in range(2):
        e = sy.host_uniform(N, nc, 70 + 10 * kind + step, -XTAL_AMPL[kind], XTAL_AMPL[kind])
        r = oracle.wrapped_crystal(kind, p, e, st)
        sig, tan, so = np.zeros((nc, N)), np.zeros((nc * nc, N)), np.zeros((42, N))
        act, it, oi = np.zeros(N, np.uint16), np.zeros(N, np.uint16), np.zeros(N, np.uint8)
        assert tw.twin_wrapped_crystal(kind, C.byref(p), H.P(e), C.c_double(1.0), H.P(st), H.P(sig), H.P(tan), H.P(so), H.P(act), H.P(it), H.P(oi),
                                       C.c_double(1e-8), 100, C.c_double(1e-8), 10, C.c_int64(N)) == 0
        nch += _xtal_compare(r, act, it, oi, sig, tan, so, kind)
        good = (r["active"] & 0xA000) == 0
        st = np.where(good[None, :], r["state"], st)
    print("wrapped crystal kind %d: %d chaotic / failed of %d" % (kind, nch, 2 * N))


@pytest.mark.gpu
@pytest.mark.parametrize("lname,lay", [("soa", 0), ("aos", 1)])
@pytest.mark.parametrize("kind", KINDS)
def test_gpu_wrapped_crystal(mm, cuda_lib, oracle, kind, lname, lay):
    import torch
    from test_gpu_parity import make_crystal
    dim = {1: mm.UniaxialStrain(), 2: mm.PlaneStrain(), 3: mm.UniaxialStress(), 4: mm.PlaneStress(), 5: mm.ShellStress()}[kind]
    m = make_crystal(mm)
    p = H.crystal_c(oracle)
    N = 3001
    nc = NC[kind]
    T = (lambda a: a) if lay == 0 else (lambda a: np.ascontiguousarray(a.T))
    st = mm.initial_material_state(m, N, layout=lname)
    st_ref = np.zeros((42, N))
    for step in range(2):
        e = sy.host_uniform(N, nc, 170 + 10 * kind + step, -XTAL_AMPL[kind], XTAL_AMPL[kind])
        r = oracle.wrapped_crystal(kind, p, e, st_ref)
        s, t, new = mm.material_response(dim, m, torch.from_numpy(T(e)).cuda(), st, 1.0, options={"defer_check": True})
        g = lambda x: x.cpu().numpy() if lay == 0 else np.ascontiguousarray(x.cpu().numpy().T)
        _xtal_compare(r, new.flags.cpu().numpy(), new.iters.cpu().numpy(), new.outer_iters.cpu().numpy(), g(s), g(t), g(new.data), kind)
        # host-buffer form: the same kernels -> identical bits
        sth = mm.CrystalViscoPlasticState(st.data.cpu().numpy(), lname)
        sh, th, nh = mm.material_response(dim, m, T(e), sth, 1.0, options={"defer_check": True})
        assert np.array_equal(sh, s.cpu().numpy()) and np.array_equal(th, t.cpu().numpy()) and np.array_equal(nh.data, new.data.cpu().numpy())
        assert np.array_equal(nh.flags, new.flags.cpu().numpy().view(np.uint16)) and int(nh.n_fail[0]) == int(new.n_fail.item())
        good = (r["active"] & 0xA000) == 0
        st_ref = np.where(good[None, :], r["state"], st_ref)
        keep = torch.from_numpy(good).cuda()
        data = torch.where(keep[None, :] if lay == 0 else keep[:, None], new.data, st.data)
        st = mm.CrystalViscoPlasticState(data, lname)
        st_ref_dev = g(data)
        assert H.rel_err(st_ref_dev[:, good], st_ref[:, good]).max() < H.RTOL
        st_ref = st_ref_dev                                                    # keep both sides on the same state bits


@pytest.mark.gpu
@pytest.mark.parametrize("lname,lay", [("soa", 0), ("aos", 1)])
@pytest.mark.parametrize("kind", [3, 4, 5])
def test_gpu_wrapped_crystal_rounds_equal_thread_per_point(mm, cuda_lib, kind, lname, lay):
    """Stress-restricted kinds around the crystal: the rounds of the two-kernel crystal path (default) against the one-thread-per-point
    kernel (mm_set_tuning("wrapped_crystal_variant", 0)) — masks, inner and outer counts identical on every regular point, values within
    1e-12; also through several workspace slabs, without a tangent, and with the state updated in place through the C ABI."""
    import torch
    from test_gpu_parity import make_crystal
    dim = {3: mm.UniaxialStress(), 4: mm.PlaneStress(), 5: mm.ShellStress()}[kind]
    m = make_crystal(mm)
    N = 5003
    nc = NC[kind]
    T = (lambda a: a) if lay == 0 else (lambda a: np.ascontiguousarray(a.T))
    g = lambda x: x.cpu().numpy() if lay == 0 else np.ascontiguousarray(x.cpu().numpy().T)

    def run(variant, slab=0, tangent=True, inplace=False):
        olds = [(b"wrapped_crystal_variant", cuda_lib.mm_set_tuning(b"wrapped_crystal_variant", variant)),
                (b"wrapped_crystal_slab", cuda_lib.mm_set_tuning(b"wrapped_crystal_slab", slab))]
        try:
            st = mm.initial_material_state(m, N, layout=lname)
            outs = []
            for step in range(2):
                e = torch.from_numpy(T(sy.host_uniform(N, nc, 270 + 10 * kind + step, -XTAL_AMPL[kind], XTAL_AMPL[kind]))).cuda()
                if inplace:
                    P = lambda t_: C.c_void_p(t_.data_ptr())
                    shp = (lambda n_: (n_, N)) if lay == 0 else (lambda n_: (N, n_))
                    s = torch.empty(shp(nc), dtype=torch.float64, device="cuda")
                    t = torch.empty(shp(nc * nc), dtype=torch.float64, device="cuda")
                    ac, it_ = torch.empty(N, dtype=torch.int16, device="cuda"), torch.empty(N, dtype=torch.int16, device="cuda")
                    oi, nf = torch.empty(N, dtype=torch.uint8, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
                    data = st.data.clone()
                    rc = cuda_lib.mm_wrapped_crystal(kind, C.byref(m._c), P(e), C.c_double(1.0), P(data), P(s), P(t), P(data), P(ac), P(it_), P(oi), P(nf), None, None, N, lay, None)
                    assert rc == 0
                    torch.cuda.synchronize()
                    outs.append((g(s), g(t), g(data), ac.cpu().numpy().view(np.uint16), it_.cpu().numpy().view(np.uint16), oi.cpu().numpy()))
                    st = mm.CrystalViscoPlasticState(data, lname)
                    continue
                s, t, new = mm.material_response(dim, m, e, st, 1.0, tangent=tangent, options={"defer_check": True})
                outs.append((g(s), None if t is None else g(t), g(new.data), new.flags.cpu().numpy().view(np.uint16), new.iters.cpu().numpy().view(np.uint16),
                             new.outer_iters.cpu().numpy()))
                # both variants continue from the SAME state bits: the thread-per-point result is the one carried over
                st = new
            return outs
        finally:
            for name, v in olds:
                cuda_lib.mm_set_tuning(name, v)

    def same(a, b, first_step_only=False):
        for k_, ((s1, t1, d1, f1, i1, o1), (s2, t2, d2, f2, i2, o2)) in enumerate(zip(a, b)):
            if first_step_only and k_ > 0:
                break
            ok = ((f1 & 0xA000) == 0) & (i1 < H.CHAOTIC_ITERS) & ((f2 & 0xA000) == 0) & (i2 < H.CHAOTIC_ITERS)
            assert ok.mean() > 0.97
            assert np.array_equal(f1[ok], f2[ok]) and np.array_equal(i1[ok], i2[ok]) and np.array_equal(o1[ok], o2[ok])
            assert sig_err(s1[:, ok], s2[:, ok]) < H.RTOL and H.rel_err(d1[:, ok], d2[:, ok]).max() < H.RTOL
            if t1 is not None:
                assert H.rel_err(t1[:, ok], t2[:, ok]).max() < max(TAN_TOL[kind], 1e-10)

    ref = run(0)
    assert ((ref[0][3] & 0xFFF) != 0).mean() > 0.2
    new = run(1)
    same(ref, new, first_step_only=True)          # step 2 starts from each variant's own step-1 state (equal to 1e-12, not bitwise)
    sl = run(1, slab=1000)
    for x, y in zip(new, sl):                      # slabs only change where the batch lives: identical bits
        assert all(np.array_equal(p_, q_) for p_, q_ in zip(x, y) if p_ is not None)
    nt = run(1, tangent=False)
    for x, y in zip(new, nt):
        assert np.array_equal(x[0], y[0]) and np.array_equal(x[2], y[2]) and np.array_equal(x[3], y[3])
    ip = run(1, inplace=True)
    for x, y in zip(new, ip):
        assert all(np.array_equal(p_, q_) for p_, q_ in zip(x, y))
