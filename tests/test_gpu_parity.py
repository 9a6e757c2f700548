"""GPU parity: the CUDA path (through the C ABI, via the host mirror) against the oracle on the same
seeded inputs, against the reference's golden vectors, and — at BASELINE.json's full sizes — through
size-independent properties plus an oracle check on a strided sample.

Bars (BASELINE.json north_star): yield / active-slip decisions and Newton iteration counts exact;
σ, dσdε, state within 1e-12 relative (FP64).  "Relative" is the norm-wise per-point metric of Julia's
`≈` (what the reference's own tests use); the Plastic state is scaled by its conjugate total quantity
(see helpers.plastic_state_scales).
"""
import numpy as np
import pytest

import helpers as H
import materialmodels_jl_b200._abi as abi
import materialmodels_jl_b200.synth as sy

pytestmark = pytest.mark.gpu

LAYOUTS = [("soa", 0), ("aos", 1)]


@pytest.fixture(scope="module")
def torch(cuda_lib):
    import torch

    assert torch.cuda.is_available()
    return torch


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def to_layout(a_soa, lay):
    return a_soa if lay == 0 else np.ascontiguousarray(a_soa.T)


def as_soa(a, lay):
    return a if lay == 0 else np.ascontiguousarray(a.T)


# ---------------------------------------------------------------------------------------------
def test_device_generator_bit_identical(mm, torch):
    for lname, lay in LAYOUTS:
        for nc, lo, hi, seed in (sy.c1_strain(), sy.c2_strain(1), sy.c3_defgrad(), sy.c4_strain_increment(0), sy.c5_jump()):
            d = host(mm.synth_uniform(100003, nc, seed, lo, hi, layout=lname))
            h = sy.host_uniform(100003, nc, seed, lo, hi, layout=lay)
            assert np.array_equal(d, h)


@pytest.mark.parametrize("lname,lay", LAYOUTS)
@pytest.mark.parametrize("N", [1, 255, 256, 257, 100003])
def test_linear_elastic(mm, torch, oracle, lname, lay, N):
    m = mm.LinearElastic(E=200e3, ν=0.3)
    nc, lo, hi, seed = sy.c1_strain()
    eps = sy.host_uniform(N, nc, seed, lo, hi, layout=lay)
    s_ref, t_ref = oracle.linear_elastic(200e3, 0.3, eps, layout=lay)
    s, t, st = mm.material_response(m, dev(torch, eps), layout=lname)
    assert H.rel_err(as_soa(host(s), lay), as_soa(s_ref, lay)).max() < H.RTOL
    assert np.array_equal(host(t), t_ref)                                  # tangent == Eᵉ exactly (test_linear_elastic.jl:18)
    s2, t2, _ = mm.material_response(m, dev(torch, eps), layout=lname, tangent=False)
    assert np.array_equal(host(s2), host(s)) and np.array_equal(t2, oracle.elastic_tangent_3d(200e3, 0.3))


def test_empty_batches(mm, torch):
    m = mm.LinearElastic(E=200e3, ν=0.3)
    s, t, _ = mm.material_response(m, torch.zeros((6, 0), dtype=torch.float64, device="cuda"))
    assert s.shape == (6, 0)
    p = mm.Plastic(**sy.PLASTIC_PARAMS)
    st = mm.initial_material_state(p, 0)
    s, t, st2 = mm.material_response(p, torch.zeros((6, 0), dtype=torch.float64, device="cuda"), st)
    assert s.shape == (6, 0) and st2.data.shape == (14, 0)
    z = lambda nc: torch.zeros((nc, 0), dtype=torch.float64, device="cuda")
    # every other entry point accepts an empty batch too (N = 0 is a no-op, never a launch with a zero grid)
    s, t, _ = mm.material_response(mm.dPdF(), mm.NeoHook(λ=1.0, μ=1.0), mm.DeformationGradient(z(9)))
    assert s.shape == (9, 0) and t.shape == (81, 0)
    assert mm.elastic_strain_energy_density(mm.Yeoh(λ=1.0, μ=1.0, c2=-1.0, c3=1.0), mm.GreenLagrange(z(6))).shape == (0,)
    x = sy.XU_PARAMS
    xm = mm.XuNeedleman(σₘₐₓ=x["sigma_max"], τₘₐₓ=x["tau_max"], Φₙ=mm.xu_needleman_Φₙ(x["sigma_max"], x["delta_n"]),
                        Φₜ=mm.xu_needleman_Φₜ(x["tau_max"], x["delta_t"]), Δₙˢ=x["Delta_n_star"])
    assert mm.material_response(xm, z(3))[1].shape == (9, 0)
    s, t, _ = mm.material_response(mm.PlaneStress(), p, z(3), mm.initial_material_state(p, 0))
    assert s.shape == (3, 0) and t.shape == (9, 0)
    ti = mm.TransverselyIsotropic(E_L=100e3, E_T=10e3, G_LT=5e3, ν_LT=0.4, ν_TT=0.3)
    s, t, _ = mm.material_response(ti, z(6), mm.TransverselyIsotropicState(z(3), "soa"))
    assert s.shape == (6, 0)
    cm = make_crystal(mm)
    s, t, cs = mm.material_response(cm, z(6), mm.initial_material_state(cm, 0), 1.0)
    assert s.shape == (6, 0) and cs.data.shape == (42, 0)
    s, t, cs = mm.material_response(mm.PlaneStress(), cm, z(3), mm.initial_material_state(cm, 0), 1.0)
    assert s.shape == (3, 0) and t.shape == (9, 0) and cs.data.shape == (42, 0)
    for lay in ("soa", "aos"):
        e = torch.zeros((6, 0) if lay == "soa" else (0, 6), dtype=torch.float64, device="cuda")
        assert mm.material_response(m, e, layout=lay)[0].numel() == 0


def test_linear_elastic_golden(mm, torch, golden):
    m = mm.LinearElastic(E=200e3, ν=0.3)
    g = golden["LinearElastic1"]
    for k, eps in enumerate(H.linear_elastic_golden_loading()):
        s, t, _ = mm.material_response(m, dev(torch, eps.reshape(6, 1)))
        assert np.allclose(host(s)[:, 0], g["stresses"][k], rtol=1e-15, atol=0)
        assert np.array_equal(host(t)[:, 0], np.array(g["tangents"][k]))


# ---------------------------------------------------------------------------------------------
def check_plastic(r, s, t, new, eps_soa, lay):
    sig, tan, so = as_soa(host(s), lay), as_soa(host(t), lay), as_soa(host(new.data), lay)
    rs, rt, rst = as_soa(r["sig"], lay), as_soa(r["tan"], lay), as_soa(r["state"], lay)
    assert np.array_equal(host(new.flags), r["flags"]), "yield flags differ"
    assert np.array_equal(host(new.iters), r["iters"]), "Newton iteration counts differ"
    assert H.rel_err(sig, rs).max() < H.RTOL
    assert H.rel_err(tan, rt).max() < H.RTOL
    es, ss = H.plastic_state_scales(eps_soa, rs)
    assert H.scaled_err(so[0:6], rst[0:6], es).max() < H.RTOL
    assert H.scaled_err(so[13:14], rst[13:14], es).max() < H.RTOL
    assert H.scaled_err(so[6:13], rst[6:13], ss).max() < H.RTOL
    H.report_state_own_norm(so, rst, r["flags"], "%d points" % so.shape[1])      # what the scaled metric waives, printed beside it


@pytest.mark.parametrize("lname,lay", LAYOUTS)
def test_plastic_parity_three_steps(mm, torch, oracle, lname, lay):
    N = 200003
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    p = H.plastic_c()
    st = mm.initial_material_state(m, N, layout=lname)
    st_ref = np.zeros((14, N) if lay == 0 else (N, 14))
    for step in range(3):
        nc, lo, hi, seed = sy.c2_strain(step)
        eps = sy.host_uniform(N, nc, seed, lo, hi, layout=lay)
        r = oracle.plastic(p, eps, st_ref, layout=lay)
        assert r["n_fail"] == 0 and 0.3 < r["flags"].mean() < 0.7
        s, t, new = mm.material_response(m, dev(torch, eps), st, cache=mm.get_cache(m))
        check_plastic(r, s, t, new, as_soa(eps, lay), lay)
        el = r["flags"] == 0
        assert np.array_equal(as_soa(host(new.data), lay)[:, el], as_soa(host(st.data), lay)[:, el])   # Plastic.jl:135
        st, st_ref = new, r["state"]


def test_plastic_golden_path(mm, torch, golden):
    """Plastic1.jld2 through the CUDA path (test_plastic.jl:33-49), 37 identical points per step."""
    g = golden["Plastic1"]
    m = mm.Plastic(E=200e3, ν=0.3, σ_y=200.0, H=50.0, r=0.5, κ_inf=13.0, α_inf=13.0)
    n = 37
    st = mm.initial_material_state(m, n)
    iters = []
    for k, eps in enumerate(H.plastic_golden_loading()):
        s, t, st = mm.material_response(m, dev(torch, np.repeat(eps.reshape(6, 1), n, axis=1)), st)
        s, t = host(s), host(t)
        assert np.all(s == s[:, :1]) and np.all(t == t[:, :1])            # deterministic across threads
        assert H.rel_err(s[:, :1], np.array(g["stresses"][k]).reshape(6, 1))[0] < 5e-15
        assert H.rel_err(t[:, :1], np.array(g["tangents"][k]).reshape(36, 1))[0] < 5e-15
        iters.append(int(host(st.iters)[0]))
    assert iters == [0, 2, 2, 1, 0, 0, 2, 2, 0, 2, 2, 1]


def test_plastic_without_tangent_and_in_place_state(mm, torch, cuda_lib):
    import ctypes as C

    N = 50001
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    nc, lo, hi, seed = sy.c2_strain(0)
    eps = mm.synth_uniform(N, nc, seed, lo, hi)
    st0 = mm.initial_material_state(m, N)
    s, t, st1 = mm.material_response(m, eps, st0)
    s2, t2, st1b = mm.material_response(m, eps, st0, tangent=False)
    assert t2 is None and torch.equal(s, s2) and torch.equal(st1.data, st1b.data)
    # state_in may alias state_out (include/mmb200.h)
    buf = st0.data.clone()
    sig = torch.empty_like(s)
    rc = cuda_lib.mm_plastic(C.byref(m._c), C.c_void_p(eps.data_ptr()), C.c_void_p(buf.data_ptr()), C.c_void_p(sig.data_ptr()), None,
                             C.c_void_p(buf.data_ptr()), None, None, None, None, N, 0, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(buf, st1.data) and torch.equal(sig, s)


@pytest.mark.parametrize("N", [1, 127, 4097, 28416, 28417, 100003])
def test_plastic_l2_prefetch_hint_changes_no_bit(mm, torch, cuda_lib, N):
    """mm_set_tuning("plastic_soa_prefetch", d): lanes 0 and 16 of every warp ask L2 for the input lines of the points d ahead (mm_ops.cu).  A
    hint: with it (default distance, a short one that stays inside small batches, one past the end) and without it every output word is the same."""
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    nc, lo, hi, seed = sy.c2_strain(1)
    nc0, lo0, hi0, seed0 = sy.c2_strain(0)
    eps = mm.synth_uniform(N, nc, seed, lo, hi)
    st0 = mm.material_response(m, mm.synth_uniform(N, nc0, seed0, lo0, hi0), mm.initial_material_state(m, N), tangent=False)[2]
    res = {}
    for d in (0, 28416, 32, N, 10**9):
        old = cuda_lib.mm_set_tuning(b"plastic_soa_prefetch", d)
        try:
            s, t, st = mm.material_response(m, eps, st0)
            torch.cuda.synchronize()
            res[d] = [x.clone() for x in (s, t, st.data, st.flags, st.iters)]
        finally:
            cuda_lib.mm_set_tuning(b"plastic_soa_prefetch", old)
    for d, got in res.items():
        for a, b in zip(res[0], got):
            assert torch.equal(a.view(torch.int64) if a.dtype == torch.float64 else a, b.view(torch.int64) if b.dtype == torch.float64 else b), d


@pytest.mark.parametrize("N", [1, 255, 65537, 200003])
def test_streaming_kernels_l2_prefetch_hint_changes_no_bit(mm, torch, cuda_lib, N):
    """mm_set_tuning("stream_soa_prefetch", d) for the LinearElastic / hyperelastic / XuNeedleman SoA kernels: -1 = their own defaults
    (tangent-writing launches only), 0 = no hint, d > 0 = that distance for every launch.  Same output words whatever the distance."""
    le = mm.LinearElastic(E=200e3, ν=0.3)
    nc, lo, hi, seed = sy.c1_strain()
    eps = mm.synth_uniform(N, nc, seed, lo, hi)
    nc, lo, hi, seed = sy.c3_defgrad()
    F = mm.synth_uniform(N, nc, seed, lo, hi)
    prm = sy.YEOH_PARAMS
    hm = mm.Yeoh(λ=prm["lambda_"], μ=prm["mu"], c2=prm["c2"], c3=prm["c3"])
    x = sy.XU_PARAMS
    xm = mm.XuNeedleman(σₘₐₓ=x["sigma_max"], τₘₐₓ=x["tau_max"], Φₙ=mm.xu_needleman_Φₙ(x["sigma_max"], x["delta_n"]),
                        Φₜ=mm.xu_needleman_Φₜ(x["tau_max"], x["delta_t"]), Δₙˢ=x["Delta_n_star"])
    nc, lo, hi, seed = sy.c5_jump()
    D = mm.synth_uniform(N, nc, seed, lo, hi)
    calls = [lambda: mm.material_response(le, eps), lambda: mm.material_response(le, eps, tangent=False),
             lambda: mm.material_response(mm.dSdC(), hm, mm.DeformationGradient(F)), lambda: mm.material_response(mm.dPdF(), hm, mm.DeformationGradient(F)),
             lambda: mm.material_response(xm, D), lambda: mm.material_response(xm, D, tangent=False)]
    res = {}
    for d in (0, -1, 32, N, 10**9):
        old = cuda_lib.mm_set_tuning(b"stream_soa_prefetch", d)
        try:
            res[d] = [[t.clone() for t in c() if isinstance(t, torch.Tensor)] for c in calls]
            torch.cuda.synchronize()
        finally:
            cuda_lib.mm_set_tuning(b"stream_soa_prefetch", old)
    for d, got in res.items():
        for ra, rb in zip(res[0], got):
            assert len(ra) == len(rb) and len(ra) >= 1
            for a, b in zip(ra, rb):
                assert torch.equal(a.view(torch.int64), b.view(torch.int64)), d


def test_plastic_nonconvergence_is_flagged_and_raises(mm, torch, oracle):
    N = 5000
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    nc, lo, hi, seed = sy.c2_strain(0)
    eps = sy.host_uniform(N, nc, seed, lo, hi)
    opts = {"nlsolve_params": {"iterations": 1}}                          # one Newton update is not enough from the trial state
    r = oracle.plastic(H.plastic_c(), eps, np.zeros((14, N)), max_iter=1)
    assert r["n_fail"] > 0
    with pytest.raises(RuntimeError, match="Material model not converged"):   # Plastic.jl:157
        mm.material_response(m, dev(torch, eps), mm.initial_material_state(m, N), options=opts)
    o2 = dict(opts, defer_check=True)
    s, t, new = mm.material_response(m, dev(torch, eps), mm.initial_material_state(m, N), options=o2)
    assert int(new.n_fail.item()) == r["n_fail"]
    assert np.array_equal(host(new.flags), r["flags"])


# ---------------------------------------------------------------------------------------------
HYPER = [("NeoHook", abi.MM_NEOHOOK, sy.NEOHOOK_PARAMS), ("Yeoh", abi.MM_YEOH, sy.YEOH_PARAMS), ("StVenant", abi.MM_STVENANT, sy.NEOHOOK_PARAMS),
         ("NeoHook", abi.MM_NEOHOOK, dict(lambda_=115384.6, mu=76923.1, c2=0.0, c3=0.0))]


@pytest.mark.parametrize("lname,lay", LAYOUTS)
@pytest.mark.parametrize("name,model,prm", HYPER)
def test_hyper_from_F_and_C(mm, torch, oracle, lname, lay, name, model, prm):
    N = 100003
    m = getattr(mm, name)(λ=prm["lambda_"], μ=prm["mu"], c2=prm["c2"], c3=prm["c3"])
    nc, lo, hi, seed = sy.c3_defgrad()
    F = sy.host_uniform(N, nc, seed, lo, hi, layout=lay)
    S_ref, T_ref = oracle.hyper(model, abi.MMHyper(**prm), F, abi.MM_STRAIN_F, layout=lay)
    S, T, _ = mm.material_response(mm.dSdC(), m, mm.DeformationGradient(dev(torch, F)), None, 0.0, layout=lname)   # traits.jl:91-107
    assert H.rel_err(as_soa(host(S), lay), as_soa(S_ref, lay)).max() < H.RTOL
    assert H.rel_err(as_soa(host(T), lay), as_soa(T_ref, lay)).max() < H.RTOL
    # native call with C (neohook.jl:35) gives the same as the F route
    Fs = as_soa(F, lay)
    C6 = np.stack([(Fs[0 + 3 * a] * Fs[0 + 3 * b] + Fs[1 + 3 * a] * Fs[1 + 3 * b] + Fs[2 + 3 * a] * Fs[2 + 3 * b]) for a, b in ((0, 0), (1, 0), (2, 0), (1, 1), (2, 1), (2, 2))])
    S2, T2, _ = mm.material_response(m, dev(torch, to_layout(C6, lay)), layout=lname)
    assert H.rel_err(as_soa(host(S2), lay), as_soa(S_ref, lay)).max() < 1e-11
    assert H.rel_err(as_soa(host(T2), lay), as_soa(T_ref, lay)).max() < 1e-11
    # S = 2 dΨ/dC, dSdC = 2 d²Ψ/dC² at a few points (test_neohook.jl:13-16), via the oracle's nested duals
    for j in (0, N // 2, N - 1):
        _, S_ad, T_ad = oracle.hyper_ad(model, abi.MMHyper(**prm), np.ascontiguousarray(C6[:, j]))
        assert H.rel_err(as_soa(host(S), lay)[:, j:j + 1], S_ad.reshape(6, 1))[0] < 1e-10
        assert H.rel_err(as_soa(host(T), lay)[:, j:j + 1], T_ad.reshape(36, 1))[0] < 1e-10


@pytest.mark.parametrize("lname,lay", LAYOUTS)
@pytest.mark.parametrize("dim", [3, 2])
def test_xu_needleman(mm, torch, oracle, lname, lay, dim):
    N = 100003
    x = sy.XU_PARAMS
    m = mm.XuNeedleman(σₘₐₓ=x["sigma_max"], τₘₐₓ=x["tau_max"], Φₙ=mm.xu_needleman_Φₙ(x["sigma_max"], x["delta_n"]),
                       Φₜ=mm.xu_needleman_Φₜ(x["tau_max"], x["delta_t"]), Δₙˢ=x["Delta_n_star"])
    nc, lo, hi, seed = sy.c5_jump()
    D = sy.host_uniform(N, nc, seed, lo, hi)
    if dim == 2:
        D = np.ascontiguousarray(D[1:])
    D = to_layout(D, lay)
    T_ref, H_ref = oracle.xu_needleman(H.xu_c(oracle), D, dim=dim, layout=lay)
    T, Hh, _ = mm.material_response(m, dev(torch, D), layout=lname)
    assert H.rel_err(as_soa(host(T), lay), as_soa(T_ref, lay)).max() < H.RTOL
    assert H.rel_err(as_soa(host(Hh), lay), as_soa(H_ref, lay)).max() < H.RTOL
    # KATs of test_xu_needleman.jl:15-28 through the CUDA path
    kat = np.zeros((dim, 2))
    kat[dim - 1, 0] = 0.002
    kat[0, 1] = np.sqrt(2) * 0.002 / 2
    Tk, _, _ = mm.material_response(m, dev(torch, to_layout(kat, lay)), layout=lname)
    Tk = as_soa(host(Tk), lay)
    assert np.isclose(Tk[dim - 1, 0], 400.0) and np.isclose(Tk[0, 1], 200.0)


@pytest.mark.parametrize("dim", [3, 2])
def test_xu_needleman_soa_tile_kernel_equals_direct_kernel(mm, torch, cuda_lib, dim):
    """The copy-engine tile kernel for SoA arrays (mm_set_tuning("xu_soa_tiled", 1)) against the direct kernel: identical bits, for full
    tiles, a ragged last tile, an odd point count (component runs not 16-byte aligned -> element-wise copies) and without the Hessian."""
    x = sy.XU_PARAMS
    m = mm.XuNeedleman(σₘₐₓ=x["sigma_max"], τₘₐₓ=x["tau_max"], Φₙ=mm.xu_needleman_Φₙ(x["sigma_max"], x["delta_n"]),
                       Φₜ=mm.xu_needleman_Φₜ(x["tau_max"], x["delta_t"]), Δₙˢ=x["Delta_n_star"])
    nc, lo, hi, seed = sy.c5_jump()
    for N in (1, 255, 256, 100000, 100002, 100003, 1 << 20):
        D = sy.host_uniform(N, nc, seed, lo, hi)
        if dim == 2:
            D = np.ascontiguousarray(D[1:])
        Dd = dev(torch, D)
        res = {}
        for tiled in (0, 1):
            old = cuda_lib.mm_set_tuning(b"xu_soa_tiled", tiled)
            try:
                T, Hh, _ = mm.material_response(m, Dd)
                T2, none, _ = mm.material_response(m, Dd, tangent=False)
                res[tiled] = (host(T), host(Hh), host(T2))
                assert none is None
            finally:
                cuda_lib.mm_set_tuning(b"xu_soa_tiled", old)
        for a, b in zip(res[0], res[1]):
            assert np.array_equal(a, b)


# ---------------------------------------------------------------------------------------------
def make_crystal(mm, q=0.0, structure=None):
    prm = dict(sy.CRYSTAL_PARAMS)
    prm["q"] = q
    ss = mm.slipsystems(structure or mm.FCC(), sy.CRYSTAL_RODRIGUES)
    return mm.CrystalViscoPlastic(slipsystems=ss, **prm)


@pytest.mark.parametrize("lname,lay", LAYOUTS)
@pytest.mark.parametrize("force_general", [False, True])
def test_crystal_parity_three_steps(mm, torch, oracle, lname, lay, force_general, cuda_lib):
    """FCC, q = 0 (BASELINE.json configs[3]) through both device paths: structured register-resident kernel (default)
    and the general pivoted-LU kernel (mm_set_tuning("crystal_force_general", 1))."""
    old = cuda_lib.mm_set_tuning(b"crystal_force_general", 1 if force_general else 0)
    try:
        _crystal_three_steps(mm, torch, oracle, lname, lay)
    finally:
        cuda_lib.mm_set_tuning(b"crystal_force_general", old)


def _crystal_three_steps(mm, torch, oracle, lname, lay):
    N = 20011
    m = make_crystal(mm)
    cp = H.crystal_c(oracle)
    st = mm.initial_material_state(m, N, layout=lname)
    st_ref = np.zeros((42, N) if lay == 0 else (N, 42))
    for step in range(3):
        nc, lo, hi, seed = sy.c4_strain_increment(step)
        de = sy.host_uniform(N, nc, seed, lo, hi, layout=lay)
        r = oracle.crystal(cp, de, st_ref, layout=lay)
        s, t, new = mm.material_response(m, dev(torch, de), st, 1.0, options={"defer_check": True})
        rs = dict(r, sig=as_soa(r["sig"], lay), tan=as_soa(r["tan"], lay), state=as_soa(r["state"], lay))
        n_chaotic = H.crystal_compare(rs, host(new.flags), host(new.iters), as_soa(host(s), lay), as_soa(host(t), lay), as_soa(host(new.data), lay),
                                      nfail_gpu=int(new.n_fail.item()))
        print("crystal FCC q=0 step %d: %d chaotic points of %d (oracle max iterations %d)" % (step, n_chaotic, N, int(r["iters"].max())))
        if step < 2:
            assert n_chaotic == 0        # steps A and B: masks and Newton counts EXACT on every point, the exemption is not used
        ok = (r["active"] & abi.MM_ACTIVE_FAILED) == 0
        st_ref = np.where(ok[None, :] if lay == 0 else ok[:, None], r["state"], st_ref)
        st = mm.CrystalViscoPlasticState(dev(torch, st_ref), lname)


@pytest.mark.parametrize("variant,lname,lay", [(0, "soa", 0), (1, "soa", 0), (2, "soa", 0), (3, "soa", 0), (4, "soa", 0), (5, "soa", 0), (6, "soa", 0), (3, "aos", 1), (1, "aos", 1), (5, "aos", 1), (0, "aos", 1)])
def test_crystal_newton_kernel_variants(mm, torch, oracle, cuda_lib, variant, lname, lay):
    """The Newton passes of the structured path — dense branch-free, sparse per-lane lists, phased, the sparse / dense pass over chunks
    re-ordered by first-pass candidate count (3 / 4; a pure scheduling change), and the active-set pass with / without the re-ordering
    (5 = the default / 6) — must all reproduce the oracle's masks and iteration counts exactly and its values within 1e-12
    (mm_set_tuning("crystal_variant", v))."""
    old = cuda_lib.mm_set_tuning(b"crystal_variant", variant)
    try:
        _crystal_three_steps(mm, torch, oracle, lname, lay)
    finally:
        cuda_lib.mm_set_tuning(b"crystal_variant", old)


@pytest.mark.parametrize("chunk", [64, 96, 1024])
def test_crystal_presorted_chunks_visit_every_point_once(mm, torch, cuda_lib, chunk):
    """The chunk-local re-ordering is a permutation: with it (variants 3, 5) every point gets exactly the result of the natural order
    (variants 1, 6), bit for bit, whatever the chunk size and for a ragged last chunk."""
    N = 50021
    m = make_crystal(mm)
    res = {}
    oldc = cuda_lib.mm_set_tuning(b"crystal_chunk", chunk)
    oldv = cuda_lib.mm_set_tuning(b"crystal_variant", 1)
    try:
        for variant in (1, 3, 6, 5):
            cuda_lib.mm_set_tuning(b"crystal_variant", variant)
            st = mm.initial_material_state(m, N)
            for step in range(2):
                nc, lo, hi, seed = sy.c4_strain_increment(step)
                de = mm.synth_uniform(N, nc, seed, lo, hi)
                s, t, st = mm.material_response(m, de, st, 1.0, options={"defer_check": True})
            res[variant] = (s, t, st.data, st.flags, st.iters)
    finally:
        cuda_lib.mm_set_tuning(b"crystal_variant", oldv)
        cuda_lib.mm_set_tuning(b"crystal_chunk", oldc)
    for a, b in zip(res[1], res[3]):
        assert torch.equal(a, b)
    for a, b in zip(res[6], res[5]):
        assert torch.equal(a, b)
    assert torch.equal(res[1][3], res[5][3]) and torch.equal(res[1][4], res[5][4])       # sparse vs active-set pass: same masks, same counts
    assert int((res[3][3] != 0).sum()) > N // 10          # the workload does yield


@pytest.mark.parametrize("variant", [0, 1, 5, 6])
def test_crystal_zero_sign_candidates_and_idle_points(mm, torch, oracle, cuda_lib, variant):
    """Edge inputs of the active-set launch: candidates whose frozen sign is 0 (handed to the sparse launch), points without any candidate,
    and a NaN strain increment (ends as not converged, like the oracle's) — every Newton pass variant against the oracle."""
    from test_host_logic import _crystal_edge_inputs

    de, st = _crystal_edge_inputs(4096 + 37)
    N = de.shape[1]
    m = make_crystal(mm)
    cp = H.crystal_c(oracle)
    r = oracle.crystal(cp, de, st)
    old = cuda_lib.mm_set_tuning(b"crystal_variant", variant)
    try:
        s, t, new = mm.material_response(m, dev(torch, de), mm.CrystalViscoPlasticState(dev(torch, st), "soa"), 1.0, options={"defer_check": True})
        H.crystal_compare(r, host(new.flags), host(new.iters), host(s), host(t), host(new.data), nfail_gpu=int(new.n_fail.item()))
        de2 = de.copy()
        de2[2, 5] = np.nan
        s, t, new = mm.material_response(m, dev(torch, de2), mm.CrystalViscoPlasticState(dev(torch, st), "soa"), 1.0, options={"defer_check": True})
        fl = np.asarray(host(new.flags)).view(np.uint16)
        assert (int(fl[5]) & abi.MM_ACTIVE_FAILED) and int(new.n_fail.item()) == 1
        ok = np.ones(N, bool)
        ok[5] = False
        assert np.array_equal(fl[ok], r["active"][ok]) and np.array_equal(host(new.iters)[ok], r["iters"][ok])
    finally:
        cuda_lib.mm_set_tuning(b"crystal_variant", old)


@pytest.mark.parametrize("N", [1, 4097, 50021])
def test_crystal_finish_l2_prefetch_hint_changes_no_bit(mm, torch, cuda_lib, N):
    """mm_set_tuning("crystal_finish_prefetch", d): the finish kernel asks L2 for the SoA input lines of the points d ahead (mm_crystal.cu).
    A hint: σ, tangent, state, masks and counts are the same words with it (default, a short distance, one past the end) and without it."""
    cm = make_crystal(mm)
    st = mm.initial_material_state(cm, N)
    res = {}
    for d in (0, 4096, 32, N, 10**9):
        old = cuda_lib.mm_set_tuning(b"crystal_finish_prefetch", d)
        try:
            s0 = st
            got = []
            for step in (0, 1):
                nc, lo, hi, seed = sy.c4_strain_increment(step)
                de = mm.synth_uniform(N, nc, seed, lo, hi)
                sg, tg, s0 = mm.material_response(cm, de, s0, 1.0, options={"defer_check": True})
                torch.cuda.synchronize()
                got += [sg.clone(), tg.clone(), s0.data.clone(), s0.flags.clone(), s0.iters.clone()]
            res[d] = got
        finally:
            cuda_lib.mm_set_tuning(b"crystal_finish_prefetch", old)
    for d, got in res.items():
        for a, b in zip(res[0], got):
            assert torch.equal(a.view(torch.int64) if a.dtype == torch.float64 else a, b.view(torch.int64) if b.dtype == torch.float64 else b), d


def test_crystal_golden_path(mm, torch, golden):
    g = golden["CrystalViscoPlastic1"]
    m = make_crystal(mm)
    n = 5
    st = mm.initial_material_state(m, n)
    tol_s = [1e-15, 1e-15, 5e-14, 5e-11]
    tol_t = [1e-15, 1e-15, 1e-13, 3e-10]
    iters = []
    for k, de in enumerate(H.crystal_golden_loading()):
        s, t, st = mm.material_response(m, dev(torch, np.repeat(de.reshape(6, 1), n, axis=1)), st)   # Δt defaults to 1.0
        assert H.rel_err(host(s)[:, :1], np.array(g["stresses"][k]).reshape(6, 1))[0] < tol_s[k]
        assert H.rel_err(host(t)[:, :1], np.array(g["tangents"][k]).reshape(36, 1))[0] < tol_t[k]
        iters.append(int(host(st.iters)[0]))
    assert iters == [0, 0, 19, 12]


@pytest.mark.parametrize("q,structure", [(0.4, 1), (0.3, 0), (0.0, 1)])
def test_crystal_cross_hardening_bcc_and_failures(mm, torch, oracle, q, structure):
    """Cross hardening (rank-1 term of the structured solve) and the reference's BCC table (non-traceless systems,
    fix-up pass exercised); some points exceed max_iter = 100 — flagged like the oracle's (chaotic points excepted,
    see helpers.crystal_compare), and the host mirror raises the reference's error."""
    N = 4000
    m = make_crystal(mm, q=q, structure=mm.BCC() if structure else mm.FCC())
    cp = H.crystal_c(oracle, q=q, structure=structure)
    st = mm.initial_material_state(m, N)
    st_ref = np.zeros((42, N))
    raised = False
    for step in range(3):
        nc, lo, hi, seed = sy.c4_strain_increment(step)
        de = sy.host_uniform(N, nc, seed, lo, hi)
        r = oracle.crystal(cp, de, st_ref)
        s, t, new = mm.material_response(m, dev(torch, de), st, 1.0, options={"defer_check": True})
        H.crystal_compare(r, host(new.flags), host(new.iters), host(s), host(t), host(new.data), nfail_gpu=int(new.n_fail.item()),
                          tan_tol=H.BCC_TAN_TOL if structure == 1 else H.RTOL)
        if int(new.n_fail.item()):
            with pytest.raises(RuntimeError, match="Material model not converged"):
                mm.material_response(m, dev(torch, de), st, 1.0)
            raised = True
        ok = (r["active"] & abi.MM_ACTIVE_FAILED) == 0
        st_ref = np.where(ok[None, :], r["state"], st_ref)
        st = mm.CrystalViscoPlasticState(dev(torch, st_ref), "soa")
    assert raised or structure == 0


def test_crystal_without_active_array(mm, torch, cuda_lib):
    """active == NULL: the structured path uses a stream-ordered temporary for its fix-up flags."""
    import ctypes as C

    N = 10007
    m = make_crystal(mm)
    nc, lo, hi, seed = sy.c4_strain_increment(0)
    de = mm.synth_uniform(N, nc, seed, lo, hi)
    st0 = mm.initial_material_state(m, N)
    s, t, new = mm.material_response(m, de, st0, 1.0)
    sig = torch.empty_like(s)
    so = torch.empty_like(new.data)
    rc = cuda_lib.mm_crystal(C.byref(m._c), C.c_void_p(de.data_ptr()), 1.0, C.c_void_p(st0.data.data_ptr()), C.c_void_p(sig.data_ptr()), None,
                             C.c_void_p(so.data_ptr()), None, None, None, None, N, 0, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(sig, s) and torch.equal(so, new.data)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lname,lay", LAYOUTS)
def test_host_buffer_entry_points_match_device_path(mm, torch, lname, lay):
    """mmh_* (host pointers, pipelined chunks incl. a ragged last chunk) == mm_* bit for bit."""
    import materialmodels_jl_b200._lib as L

    lib = L.load()
    old = lib.mmh_set_chunk_points(1000)
    try:
        N = 3507
        m = mm.Plastic(**sy.PLASTIC_PARAMS)
        nc, lo, hi, seed = sy.c2_strain(0)
        eps = sy.host_uniform(N, nc, seed, lo, hi, layout=lay)
        sd, td, nd = mm.material_response(m, dev(torch, eps), mm.initial_material_state(m, N, layout=lname))
        sh, th, nh = mm.material_response(m, eps, mm.initial_material_state(m, N, device="cpu", layout=lname))
        assert isinstance(sh, np.ndarray)
        assert np.array_equal(sh, host(sd)) and np.array_equal(th, host(td)) and np.array_equal(nh.data, host(nd.data))
        assert np.array_equal(nh.flags, host(nd.flags)) and np.array_equal(nh.iters, host(nd.iters))
        # second step with non-trivial state
        nc, lo, hi, seed = sy.c2_strain(1)
        eps = sy.host_uniform(N, nc, seed, lo, hi, layout=lay)
        sd, td, nd2 = mm.material_response(m, dev(torch, eps), nd)
        sh, th, nh2 = mm.material_response(m, eps, nh)
        assert np.array_equal(sh, host(sd)) and np.array_equal(th, host(td)) and np.array_equal(nh2.data, host(nd2.data))
        # the stateless models
        nc, lo, hi, seed = sy.c3_defgrad()
        F = sy.host_uniform(N, nc, seed, lo, hi, layout=lay)
        y = mm.Yeoh(λ=1.0, μ=1.0, c2=-1.0, c3=1.0)
        Sd, Td, _ = mm.material_response(mm.dSdC(), y, mm.DeformationGradient(dev(torch, F)), layout=lname)
        Sh, Th, _ = mm.material_response(mm.dSdC(), y, mm.DeformationGradient(F), layout=lname)
        assert np.array_equal(Sh, host(Sd)) and np.array_equal(Th, host(Td))
        le = mm.LinearElastic(E=200e3, ν=0.3)
        e6 = sy.host_uniform(N, 6, 1, -1e-3, 1e-3, layout=lay)
        a, b, _ = mm.material_response(le, e6, layout=lname)
        c, d, _ = mm.material_response(le, dev(torch, e6), layout=lname)
        assert np.array_equal(a, host(c)) and np.array_equal(b, host(d))
        x = sy.XU_PARAMS
        xm = mm.XuNeedleman(σₘₐₓ=x["sigma_max"], τₘₐₓ=x["tau_max"], Φₙ=mm.xu_needleman_Φₙ(x["sigma_max"], x["delta_n"]),
                            Φₜ=mm.xu_needleman_Φₜ(x["tau_max"], x["delta_t"]), Δₙˢ=x["Delta_n_star"])
        nc, lo, hi, seed = sy.c5_jump()
        D = sy.host_uniform(N, nc, seed, lo, hi, layout=lay)
        a, b, _ = mm.material_response(xm, D, layout=lname)
        c, d, _ = mm.material_response(xm, dev(torch, D), layout=lname)
        assert np.array_equal(a, host(c)) and np.array_equal(b, host(d))
        cm = make_crystal(mm)
        nc, lo, hi, seed = sy.c4_strain_increment(0)
        de = sy.host_uniform(N, nc, seed, lo, hi, layout=lay)
        a, b, sa = mm.material_response(cm, de, mm.initial_material_state(cm, N, device="cpu", layout=lname))
        c, d, sc = mm.material_response(cm, dev(torch, de), mm.initial_material_state(cm, N, layout=lname))
        assert np.array_equal(a, host(c)) and np.array_equal(b, host(d)) and np.array_equal(sa.data, host(sc.data))
        assert np.array_equal(sa.flags, host(sc.flags).view(np.uint16))
    finally:
        lib.mmh_set_chunk_points(old)


# ---------------------------------------------------------------------------------------------
def test_nan_inputs_terminate_are_flagged_and_do_not_leak(mm, torch):
    """A NaN strain can never satisfy a convergence test: the local Newton loops must still terminate (iteration caps), the point
    must be flagged as failed and counted, and its neighbours must be untouched."""
    N = 4096
    bad = torch.tensor([3, 77, 1000, 4095], device="cuda")
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    nc, lo, hi, seed = sy.c2_strain(0)
    eps = mm.synth_uniform(N, nc, seed, lo, hi)
    s0, t0, n0 = mm.material_response(m, eps, mm.initial_material_state(m, N))
    e2 = eps.clone()
    e2[0, bad] = float("nan")
    with pytest.raises(RuntimeError, match="Material model not converged"):
        mm.material_response(m, e2, mm.initial_material_state(m, N), options={"nlsolve_params": {"iterations": 50}})
    s, t, new = mm.material_response(m, e2, mm.initial_material_state(m, N), options={"nlsolve_params": {"iterations": 50}, "defer_check": True})
    assert int(new.n_fail.item()) == 4 and bool(((new.flags[bad] & abi.MM_FLAG_FAILED) != 0).all())
    ok = torch.ones(N, dtype=torch.bool, device="cuda")
    ok[bad] = False
    assert torch.equal(s[:, ok], s0[:, ok]) and torch.equal(t[:, ok], t0[:, ok]) and torch.equal(new.data[:, ok], n0.data[:, ok])
    assert int((new.flags[ok] & abi.MM_FLAG_FAILED).sum().item()) == 0
    # crystal plasticity (two-kernel structured path): same guarantees
    cm = make_crystal(mm)
    nc, lo, hi, seed = sy.c4_strain_increment(0)
    de = mm.synth_uniform(N, nc, seed, lo, hi)
    c0 = mm.material_response(cm, de, mm.initial_material_state(cm, N), 1.0)
    d2 = de.clone()
    d2[2, bad] = float("nan")
    sg, tg, cs = mm.material_response(cm, d2, mm.initial_material_state(cm, N), 1.0, options={"defer_check": True})
    fl = cs.flags.view(torch.int16).to(torch.int32) & 0xFFFF
    assert int(cs.n_fail.item()) == 4 and bool(((fl[bad] & abi.MM_ACTIVE_FAILED) != 0).all()) and int((fl[ok] & abi.MM_ACTIVE_FAILED).sum().item()) == 0
    assert torch.equal(sg[:, ok], c0[0][:, ok]) and torch.equal(cs.data[:, ok], c0[2].data[:, ok])
    # plane-stress wrapper around J2
    e3 = mm.synth_uniform(N, 3, 31, -1.2e-3, 1.2e-3)
    w0 = mm.material_response(mm.PlaneStress(), m, e3, mm.initial_material_state(m, N))
    e3b = e3.clone()
    e3b[1, bad] = float("nan")
    ws, wt, wn = mm.material_response(mm.PlaneStress(), m, e3b, mm.initial_material_state(m, N), options={"nlsolve_params": {"iterations": 50}, "defer_check": True})
    assert int(wn.n_fail.item()) == 4 and bool(((wn.flags[bad] & (abi.MM_FLAG_FAILED | abi.MM_FLAG_WRAPPER_FAILED)) != 0).all())
    assert torch.equal(ws[:, ok], w0[0][:, ok]) and torch.equal(wn.data[:, ok], w0[2].data[:, ok])


@pytest.mark.parametrize("q,structure", [(0.0, None), (0.3, None), (0.0, "bcc")])
def test_crystal_in_place_state_update(mm, torch, cuda_lib, q, structure):
    """state_out == state_in (include/mmb200.h): the two-kernel path parks μ in the state output between its kernels and hands flagged
    points to a fix-up pass that restarts from the OLD state — an in-place call must still give the out-of-place result bit for bit
    (BCC / cross-hardening runs include fix-up points)."""
    import ctypes as C

    N = 20011
    m = make_crystal(mm, q=q, structure=mm.BCC() if structure == "bcc" else None)
    st = mm.initial_material_state(m, N)
    for step in (0, 1):
        nc, lo, hi, seed = sy.c4_strain_increment(step)
        de = mm.synth_uniform(N, nc, seed, lo, hi)
        s_ref, t_ref, new = mm.material_response(m, de, st, 1.0, options={"defer_check": True})
        inplace = st.data.clone()
        sig, tan = torch.empty_like(s_ref), torch.empty_like(t_ref)
        act = torch.empty(N, dtype=torch.int16, device="cuda")
        it = torch.empty(N, dtype=torch.int16, device="cuda")
        nf = torch.zeros(1, dtype=torch.int32, device="cuda")
        P = lambda a: C.c_void_p(a.data_ptr())
        rc = cuda_lib.mm_crystal(C.byref(m._c), P(de), 1.0, P(inplace), P(sig), P(tan), P(inplace), P(act), P(it), P(nf), None, N, 0, None)
        assert rc == 0
        torch.cuda.synchronize()
        assert torch.equal(inplace, new.data) and torch.equal(sig, s_ref) and torch.equal(tan, t_ref)
        assert torch.equal(act, new.flags) and torch.equal(it, new.iters) and int(nf.item()) == int(new.n_fail.item())
        st = new


def test_aos_unaligned_bases_take_the_elementwise_tiles_and_give_the_same_bits(mm, torch):
    """AoS arrays whose base is only 8-byte aligned cannot use the bulk copy engine (16-byte rule): every tile then goes through the
    element-wise copies of the SAME kernel instantiation — including the in-place σ / state staging — and the result is
    bit-identical to the aligned call."""
    N = 10007
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    nc, lo, hi, seed = sy.c2_strain(0)
    eps = mm.synth_uniform(N, nc, seed, lo, hi, layout="aos")
    st0 = mm.initial_material_state(m, N, layout="aos")
    _, _, stA = mm.material_response(m, eps, st0)
    nc, lo, hi, seed = sy.c2_strain(1)
    eps = mm.synth_uniform(N, nc, seed, lo, hi, layout="aos")
    s, t, new = mm.material_response(m, eps, stA)

    def shifted(x):                         # same values at an address that is 8 mod 16
        buf = torch.empty(x.numel() + 1, dtype=x.dtype, device=x.device)
        v = buf[1:].view(x.shape)
        v.copy_(x)
        assert v.data_ptr() % 16 == 8
        return v

    out = {"stress": shifted(torch.zeros_like(s)), "tangent": shifted(torch.zeros_like(t)), "state": shifted(torch.zeros_like(new.data))}
    s2, t2, new2 = mm.material_response(m, shifted(eps), mm.PlasticState(shifted(stA.data), "aos"), out=out)
    assert s2.data_ptr() % 16 == 8 and new2.data.data_ptr() % 16 == 8
    assert torch.equal(s2, s) and torch.equal(t2, t) and torch.equal(new2.data, new.data)
    assert torch.equal(new2.flags, new.flags) and torch.equal(new2.iters, new.iters)
    # hyperelastic AoS too (9-in / 6 + 36-out, no aliasing)
    nh = mm.NeoHook(λ=1.0, μ=1.0)
    nc, lo, hi, seed = sy.c3_defgrad()
    F = mm.synth_uniform(N, nc, seed, lo, hi, layout="aos")
    S, T, _ = mm.material_response(mm.dSdC(), nh, mm.DeformationGradient(F), layout="aos")
    S2, T2, _ = mm.material_response(mm.dSdC(), nh, mm.DeformationGradient(shifted(F)), layout="aos")
    assert torch.equal(S2, S) and torch.equal(T2, T)


@pytest.mark.parametrize("lname", ["soa", "aos"])
def test_plastic_state_updated_in_place(mm, torch, lname):
    """`state_in` may alias `state_out` (include/mmb200.h): the AoS tile kernel prefetches the NEXT tile's inputs before it stores this
    tile's outputs and stages σ / state in the input image, so an in-place update is the case that would expose an ordering bug."""
    N = 50021
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    nc, lo, hi, seed = sy.c2_strain(0)
    eA = mm.synth_uniform(N, nc, seed, lo, hi, layout=lname)
    _, _, stA = mm.material_response(m, eA, mm.initial_material_state(m, N, layout=lname))
    nc, lo, hi, seed = sy.c2_strain(1)
    eB = mm.synth_uniform(N, nc, seed, lo, hi, layout=lname)
    s, t, new = mm.material_response(m, eB, stA)
    buf = stA.data.clone()
    s2, t2, new2 = mm.material_response(m, eB, mm.PlasticState(buf, lname), out={"state": buf})
    assert new2.data.data_ptr() == buf.data_ptr()
    assert torch.equal(s2, s) and torch.equal(t2, t) and torch.equal(buf, new.data)


@pytest.mark.parametrize("lname,lay", LAYOUTS)
def test_crystal_host_call_with_device_resident_state(mm, torch, cuda_lib, lname, lay):
    """material_response(m, numpy Δε, state on the GPU, Δt, options={"state_resident": True}) -> mmh_crystal_resident: the state array is updated in place chunk by chunk
    (AoS: where it lies; SoA: staged with device-to-device copies); σ, ∂σ∂ε, masks, counts and the state equal the device path bit for bit."""
    m = make_crystal(mm)
    N = 5003
    old = cuda_lib.mmh_set_chunk_points(1000)
    try:
        st_dev = mm.initial_material_state(m, N, layout=lname)
        st_res = mm.CrystalViscoPlasticState(st_dev.data.clone(), lname)
        for step in range(2):
            nc, lo, hi, seed = sy.c4_strain_increment(step)
            d = to_layout(sy.host_uniform(N, nc, seed, lo, hi), lay)
            s1, t1, st_dev = mm.material_response(m, dev(torch, d), st_dev, 1.0, options={"defer_check": True})
            data_before = st_res.data
            s2, t2, st_res = mm.material_response(m, d, st_res, 1.0, options={"defer_check": True, "state_resident": True})
            assert isinstance(s2, np.ndarray) and st_res.data is data_before          # host outputs, the device array updated in place
            assert np.array_equal(s2, host(s1)) and np.array_equal(t2, host(t1)) and torch.equal(st_res.data, st_dev.data)
            assert np.array_equal(st_res.flags, host(st_dev.flags).view(np.uint16)) and np.array_equal(st_res.iters, host(st_dev.iters).view(np.uint16))
            assert int(st_res.n_fail[0]) == int(st_dev.n_fail.item())
    finally:
        cuda_lib.mmh_set_chunk_points(old)
