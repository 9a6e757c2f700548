"""Host-side logic that needs no GPU: synthetic-input generator, API argument handling, sharding,
and the kernels' per-point math instantiated on the host (tests/hosttwin) against the oracle."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
import materialmodels_jl_b200._abi as abi
import materialmodels_jl_b200.synth as sy
from materialmodels_jl_b200.sharding import shard_range


def test_host_generator_bit_identical_to_oracle_generator(oracle):
    for layout in (0, 1):
        for nc, lo, hi, seed in (sy.c1_strain(), sy.c2_strain(1), sy.c3_defgrad(), sy.c5_jump()):
            a = sy.host_uniform(1000, nc, seed, lo, hi, layout)
            b = oracle.synth_uniform(1000, nc, seed, lo, hi, layout)
            assert np.array_equal(a, b)
    # i0 offsets address the same global stream (used by the per-rank shards)
    nc, lo, hi, seed = sy.c2_strain(0)
    full = sy.host_uniform(1000, nc, seed, lo, hi)
    assert np.array_equal(sy.host_uniform(300, nc, seed, lo, hi, i0=500), full[:, 500:800])


def test_shard_range_partitions_exactly():
    for N in (0, 1, 7, 10**8, 10**8 + 3):
        for world in (1, 2, 3, 8):
            r = [shard_range(N, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == N
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_material_constructors_mirror_reference(mm):
    m = mm.Plastic(E=200e3, ν=0.3, σ_y=200.0, H=50.0, r=0.5, κ_inf=13.0, α_inf=13.0)
    assert m.H_κ == -25.0 and m.H_α == -2 / 3 * 0.5 * 50.0          # Plastic.jl:31-32
    m2 = mm.Plastic(**{"E": 200e3, "ν": 0.3, "σ_y": 200.0, "H": 50.0, "r": 0.5, "κ_∞": 13.0, "α_∞": 13.0})
    assert m2._c.kappa_inf == 13.0 and m2._c.alpha_inf == 13.0
    assert mm.get_cache(mm.LinearElastic(E=1.0, ν=0.3)) is None       # test_linear_elastic.jl:11-12
    assert isinstance(mm.get_cache(m), mm.PlasticCache)
    assert mm.native_tangent_type(mm.NeoHook) is mm.dSdC and mm.native_tangent_type(mm.LinearElastic) is mm.dsigmadeps
    y = mm.Yeoh(λ=1.0, μ=1.0, c2=-1.0, c3=1.0)
    assert (y._c.lambda_, y._c.mu, y._c.c2, y._c.c3) == (1.0, 1.0, -1.0, 1.0)
    with pytest.raises(TypeError):
        mm.LinearElastic(E=1.0)


def test_api_argument_validation_without_gpu(mm):
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    st = mm.initial_material_state(m, 5, device="cpu")
    assert st.data.shape == (14, 5) and st.εᵖ.shape == (6, 5) and st.μ.shape == (1, 5)
    st_aos = mm.initial_material_state(m, 5, device="cpu", layout="aos")
    assert st_aos.data.shape == (5, 14) and st_aos.α.shape == (5, 6)
    with pytest.raises(ValueError):
        mm.material_response(m, np.zeros((6, 4)), st)                   # N mismatch
    with pytest.raises(ValueError):
        mm.material_response(m, np.zeros((5, 6)), st)                   # wrong layout for this state
    with pytest.raises(TypeError):
        mm.material_response(m, np.zeros((6, 5)))                       # state required
    with pytest.raises(TypeError, match="needs DeformationGradient"):            # traits.jl:67-72: ∂Pᵀ∂F needs F
        mm.material_response(mm.dPdF(), mm.NeoHook(λ=1.0, μ=1.0), mm.RightCauchyGreen(np.zeros((6, 1))))
    with pytest.raises(RuntimeError, match="not the native strain measure"):   # traits.jl:120
        mm.material_response(mm.dSdC(), mm.NeoHook(λ=1.0, μ=1.0), mm.SmallStrain(np.zeros((6, 1))))
    with pytest.raises(TypeError):
        mm.initial_material_state(m)                                    # batched API needs N
    # caller-provided output buffers are written through raw pointers: wrong dtype / layout is refused before the library is touched
    with pytest.raises(ValueError, match="dtype"):
        mm.material_response(m, np.zeros((6, 5)), st, out={"stress": np.zeros((6, 5), np.float32)})
    with pytest.raises(ValueError, match="contiguous"):
        mm.material_response(m, np.zeros((6, 5)), st, out={"stress": np.zeros((5, 6)).T})
    with pytest.raises(ValueError, match="dtype"):
        mm.material_response(m, np.zeros((6, 5)), st, out={"iters": np.zeros(5, np.uint8)})      # Newton counts are uint16
    # NLsolve options: only method = :newton exists on the GPU path; anything else must fail loudly (Plastic.jl:146-147 forwards them)
    for method in ("trust_region", ":anderson"):
        with pytest.raises(NotImplementedError, match="newton"):
            mm.material_response(m, np.zeros((6, 5)), st, options={"nlsolve_params": {"method": method}})
    with pytest.raises(ValueError, match="tangent must be"):
        mm.material_response(m, np.zeros((6, 5)), st, tangent="sparse")


def test_product_has_no_cpu_fallback(mm, monkeypatch):
    """If the native library is missing the product raises; it never routes to the oracle or numpy."""
    import materialmodels_jl_b200._lib as L

    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "SO_PATH", "/nonexistent/libmmb200.so")
    with pytest.raises(L.NativeLibraryMissing):
        mm.material_response(mm.LinearElastic(E=1.0, ν=0.3), np.zeros((6, 3)))
    import materialmodels_jl_b200.api as api
    src = open(api.__file__).read() + open(L.__file__).read()
    assert "oracle" not in src.replace("the oracle", "")


# ---- the kernels' math on the host vs the oracle --------------------------------------------------
N_TWIN = 20000


def test_twin_plastic_matches_oracle(oracle):
    tw = H.hosttwin()
    p = H.plastic_c()
    st = np.zeros((14, N_TWIN))
    for step in range(3):
        nc, lo, hi, seed = sy.c2_strain(step)
        eps = sy.host_uniform(N_TWIN, nc, seed, lo, hi)
        r = oracle.plastic(p, eps, st)
        sig, tan, so = np.zeros((6, N_TWIN)), np.zeros((36, N_TWIN)), np.zeros((14, N_TWIN))
        fl, it = np.zeros(N_TWIN, np.uint8), np.zeros(N_TWIN, np.uint16)
        tw.twin_plastic(C.byref(p), H.P(eps), H.P(st), H.P(sig), H.P(tan), H.P(so), H.P(fl), H.P(it), C.c_double(1e-8), 1000, C.c_int64(N_TWIN))
        assert r["n_fail"] == 0 and 0.3 < r["flags"].mean() < 0.7           # mixed elastic / plastic load
        assert np.array_equal(fl, r["flags"])                               # yield decisions exact
        assert np.array_equal(it, r["iters"])                               # Newton update counts exact
        assert H.rel_err(sig, r["sig"]).max() < H.RTOL and H.rel_err(tan, r["tan"]).max() < H.RTOL
        es, ss = H.plastic_state_scales(eps, r["sig"])
        assert H.scaled_err(so[0:6], r["state"][0:6], es).max() < H.RTOL
        assert H.scaled_err(so[13:14], r["state"][13:14], es).max() < H.RTOL
        assert H.scaled_err(so[6:13], r["state"][6:13], ss).max() < H.RTOL
        el = r["flags"] == 0
        assert np.array_equal(so[:, el], st[:, el])                         # elastic points: state untouched (Plastic.jl:135)
        st = r["state"]


def test_twin_streaming_models_match_oracle(oracle):
    tw = H.hosttwin()
    N = N_TWIN
    nc, lo, hi, seed = sy.c1_strain()
    eps = sy.host_uniform(N, nc, seed, lo, hi)
    s, t = oracle.linear_elastic(200e3, 0.3, eps)
    s2, t2 = np.zeros((6, N)), np.zeros((36, N))
    tw.twin_linear_elastic(C.byref(abi.MMLinearElastic(200e3, 0.3)), H.P(eps), H.P(s2), H.P(t2), C.c_int64(N))
    assert H.rel_err(s, s2).max() < H.RTOL and np.array_equal(t, t2)
    nc, lo, hi, seed = sy.c3_defgrad()
    F = sy.host_uniform(N, nc, seed, lo, hi)
    for model, prm in ((0, sy.NEOHOOK_PARAMS), (1, sy.YEOH_PARAMS), (2, sy.NEOHOOK_PARAMS), (0, dict(lambda_=115384.6, mu=76923.1, c2=0.0, c3=0.0))):
        hp = abi.MMHyper(**prm)
        S, T = oracle.hyper(model, hp, F, 1)
        S2, T2 = np.zeros((6, N)), np.zeros((36, N))
        tw.twin_hyper(model, C.byref(hp), H.P(F), 1, H.P(S2), H.P(T2), C.c_int64(N))
        assert H.rel_err(S, S2).max() < H.RTOL and H.rel_err(T, T2).max() < H.RTOL
    nc, lo, hi, seed = sy.c5_jump()
    D = sy.host_uniform(N, nc, seed, lo, hi)
    xp = H.xu_c(oracle)
    T, Hh = oracle.xu_needleman(xp, D)
    T2, H2 = np.zeros((3, N)), np.zeros((9, N))
    tw.twin_xu(C.byref(xp), H.P(D), H.P(T2), H.P(H2), 3, C.c_int64(N))
    assert H.rel_err(T, T2).max() < H.RTOL and H.rel_err(Hh, H2).max() < H.RTOL


def test_twin_crystal_matches_oracle(oracle):
    """Both crystal paths of the kernels (general pivoted 12x12; structured SPD/Woodbury with fix-up) on the host
    against the oracle: FCC and the reference's BCC table, without and with cross hardening, three steps."""
    tw = H.hosttwin()
    N = 4000
    n_fallback_total = 0
    for q, structure in ((0.0, 0), (0.4, 1), (0.3, 0), (0.0, 1)):
        cp = H.crystal_c(oracle, q=q, structure=structure)
        st = np.zeros((42, N))
        for step in range(3):
            nc, lo, hi, seed = sy.c4_strain_increment(step)
            de = sy.host_uniform(N, nc, seed, lo, hi)
            r = oracle.crystal(cp, de, st)
            for mode in (0, 1, 2, 3, 4):          # general pivoted / structured: dense pass, sparse pass, phased pass, active-set pass (the Newton kernel's default)
                sig, tan, so = np.zeros((6, N)), np.zeros((36, N)), np.zeros((42, N))
                ac, it = np.zeros(N, np.uint16), np.zeros(N, np.uint16)
                nfb = C.c_int64(0)
                tw.twin_crystal_mode(C.byref(cp), H.P(de), C.c_double(1.0), H.P(st), H.P(sig), H.P(tan), H.P(so), H.P(ac), H.P(it),
                                     C.c_double(1e-8), 100, C.c_int64(N), mode, C.byref(nfb))
                H.crystal_compare(r, ac, it, sig, tan, so, tan_tol=H.BCC_TAN_TOL if structure == 1 else H.RTOL)
                n_fallback_total += nfb.value
                if q == 0.0 and structure == 0 and step < 2:
                    assert nfb.value == 0          # the benchmark configuration never needs the fix-up pass
            ok = (r["active"] & abi.MM_ACTIVE_FAILED) == 0
            st = np.where(ok[None, :], r["state"], st)                      # failed points keep their old state
    assert n_fallback_total > 0                     # ... but the guard does fire on the harder configurations


def _crystal_edge_inputs(N=64):
    """Points the active-set pass must hand to the 6x6 route or end at once: a previous slip increment on a stress-free, back-stress-free state with a
    zero strain increment (the trial reduced stress of that system is exactly 0: frozen sign 0 on a candidate), zero increments on a virgin state
    (no candidate at all), ordinary loaded points in between."""
    rng = np.random.default_rng(5)
    de = rng.uniform(-sy.CRYSTAL_AMPL, sy.CRYSTAL_AMPL, (6, N))
    st = np.zeros((42, N))
    de[:, 0::4] = 0.0                           # every 4th point: no increment
    st[30 + 3, 0::8] = 1e-3                     # every 8th: mu_3 != 0 with sigma = alpha = 0  ->  s_3 = 0 and x_3 != 0
    st[30 + 7, 0::8] = -2e-4
    return de, st


def test_twin_crystal_zero_sign_candidates_and_idle_points(oracle):
    tw = H.hosttwin()
    cp = H.crystal_c(oracle)
    de, st = _crystal_edge_inputs()
    N = de.shape[1]
    r = oracle.crystal(cp, de, st)
    for mode in (1, 2, 4):
        sig, tan, so = np.zeros((6, N)), np.zeros((36, N)), np.zeros((42, N))
        ac, it = np.zeros(N, np.uint16), np.zeros(N, np.uint16)
        nfb = C.c_int64(0)
        tw.twin_crystal_mode(C.byref(cp), H.P(de), C.c_double(1.0), H.P(st), H.P(sig), H.P(tan), H.P(so), H.P(ac), H.P(it),
                             C.c_double(1e-8), 100, C.c_int64(N), mode, C.byref(nfb))
        H.crystal_compare(r, ac, it, sig, tan, so)


def test_share_host_cores_sets_the_library_thread_count():
    """One process per GPU shares the host cores for the library's host-side expansion (sharding.share_host_cores -> mmh_set_host_threads)."""
    import materialmodels_jl_b200._lib as L
    from materialmodels_jl_b200 import sharding

    lib = L.load()
    try:
        n1 = sharding.share_host_cores(1)
        n8 = sharding.share_host_cores(8)
        assert n1 >= n8 >= 2
        assert lib.mmh_set_host_threads(0) == n8          # returns the previous setting
    finally:
        lib.mmh_set_host_threads(0)
