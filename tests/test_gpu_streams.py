"""Stream semantics of the device entry points: every `mm_*` call only enqueues work on the caller's stream (no host
synchronisation, no allocation on the hot path), so a time-step loop can be captured in a CUDA graph and replayed, calls on
different streams / from different host threads may overlap, and results never depend on which stream ran them."""
import threading

import numpy as np
import pytest

import materialmodels_jl_b200.synth as sy

pytestmark = pytest.mark.gpu
DEFER = {"defer_check": True}


@pytest.fixture(scope="module")
def torch(cuda_lib):
    import torch

    assert torch.cuda.is_available()
    return torch


def make_crystal(mm):
    return mm.CrystalViscoPlastic(slipsystems=mm.slipsystems(mm.FCC(), sy.CRYSTAL_RODRIGUES), **sy.CRYSTAL_PARAMS)


def _cases(mm, torch, N):
    """(name, callable) pairs; each callable runs one update on the CURRENT stream and returns the tensors it produced"""
    p = mm.Plastic(**sy.PLASTIC_PARAMS)
    nc, lo, hi, seed = sy.c2_strain(0)
    eps = mm.synth_uniform(N, nc, seed, lo, hi)
    st0 = mm.initial_material_state(p, N)
    cm = make_crystal(mm)
    nc, lo, hi, seed = sy.c4_strain_increment(0)
    de = mm.synth_uniform(N, nc, seed, lo, hi)
    cs0 = mm.initial_material_state(cm, N)
    e3 = mm.synth_uniform(N, 3, 31, -1.2e-3, 1.2e-3)
    nh = mm.Yeoh(λ=1.0, μ=1.0, c2=-1.0, c3=1.0)
    nc, lo, hi, seed = sy.c3_defgrad()
    F = mm.synth_uniform(N, nc, seed, lo, hi)

    def plastic():
        s, t, new = mm.material_response(p, eps, st0, options=DEFER)
        return [s, t, new.data, new.flags, new.iters]

    def crystal():
        s, t, new = mm.material_response(cm, de, cs0, 1.0, options=DEFER)
        return [s, t, new.data, new.flags, new.iters]

    def plane_stress():
        s, t, new = mm.material_response(mm.PlaneStress(), p, e3, st0, options=DEFER)
        return [s, t, new.data, new.flags, new.outer_iters]

    def yeoh_dpdf():
        s, t, _ = mm.material_response(mm.dPdF(), nh, mm.DeformationGradient(F))
        return [s, t]

    return [("plastic", plastic), ("crystal", crystal), ("plane_stress_plastic", plane_stress), ("yeoh_dPdF", yeoh_dpdf)]


def test_cuda_graph_capture_and_replay(mm, torch):
    """capture one update of each kind in a CUDA graph (the library must not synchronise or allocate with cudaMalloc while
    capturing), overwrite the outputs, replay, and get the eager bits back"""
    N = 30011
    for name, fn in _cases(mm, torch, N):
        ref = [x.clone() for x in fn()]            # eager; also the warm-up that does the one-time shared-memory opt-ins
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            outs = fn()
        for x in outs:
            x.fill_(0) if x.dtype != torch.float64 else x.fill_(float("nan"))
        g.replay()
        torch.cuda.synchronize()
        for a, b in zip(outs, ref):
            assert torch.equal(a, b), "%s: graph replay differs from the eager call" % name


def test_side_stream_matches_default_stream(mm, torch):
    N = 20011
    side = torch.cuda.Stream()
    for name, fn in _cases(mm, torch, N):
        ref = fn()
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            got = fn()
        side.synchronize()
        for a, b in zip(got, ref):
            assert torch.equal(a, b), name


def test_concurrent_host_threads_on_own_streams(mm, torch):
    """four host threads, each with its own stream and its own inputs, hammer the library at once (different kernels, shared
    one-time initialisation, the crystal fix-up workspace); every result must equal the single-threaded one"""
    N = 15013
    cases = _cases(mm, torch, N)
    ref = {name: [x.clone() for x in fn()] for name, fn in cases}
    torch.cuda.synchronize()
    errors = []

    def worker(name, fn):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(8):
                    got = fn()
            s.synchronize()
            for a, b in zip(got, ref[name]):
                if not torch.equal(a, b):
                    errors.append("%s: result differs under concurrency" % name)
        except Exception as e:                      # noqa: BLE001 - report, do not hang the join
            errors.append("%s: %r" % (name, e))

    th = [threading.Thread(target=worker, args=c) for c in cases]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


def test_small_batch_graph_loop_is_launch_bound_not_sync_bound(mm, torch):
    """a 20-step time loop of a small batch inside ONE graph: state chained from step to step on the device; equals the eager loop"""
    N = 4099
    p = mm.Plastic(**sy.PLASTIC_PARAMS)
    eps = [mm.synth_uniform(N, 6, 900 + k, -6.6e-4 * (1 + 0.1 * k), 6.6e-4 * (1 + 0.1 * k)) for k in range(20)]

    def loop():
        st = mm.initial_material_state(p, N)
        for e in eps:
            s, t, st = mm.material_response(p, e, st, options=DEFER)
        return s, t, st.data, st.n_fail

    ref = [x.clone() for x in loop()]
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs = loop()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    for a, b in zip(outs, ref):
        assert torch.equal(a, b)
    assert int(outs[3].item()) == 0


def test_mixed_host_buffer_entry_points_from_two_threads(mm, torch, cuda_lib):
    """mmh_* calls share ONE staging workspace (streams, device chunk buffers, fail counter) behind ONE lock: two host threads calling
    DIFFERENT entry points at the same time (ctypes releases the GIL for the call) must each get the serial result, bit for bit —
    with a small chunk size so that every call runs many pipeline chunks and the workspace is resized back and forth"""
    N1, N2 = 50001, 30011
    p = mm.Plastic(**sy.PLASTIC_PARAMS)
    nc, lo, hi, seed = sy.c2_strain(0)
    eps = sy.host_uniform(N1, nc, seed, lo, hi)
    st0 = np.zeros((14, N1))
    nh = mm.Yeoh(λ=1.0, μ=1.0, c2=-1.0, c3=1.0)
    nc, lo, hi, seed = sy.c3_defgrad()
    F = sy.host_uniform(N2, nc, seed, lo, hi)
    old = cuda_lib.mmh_set_chunk_points(2048)
    try:
        s1, t1, n1 = mm.material_response(p, eps, mm.PlasticState(st0))
        s2, t2, _ = mm.material_response(mm.dSdC(), nh, mm.DeformationGradient(F))
        s3, T3, n3 = mm.material_response(p, eps, mm.PlasticState(st0), tangent="factored")
        errs = []

        def plastic_worker():
            try:
                for _ in range(50):
                    a, b, c = mm.material_response(p, eps, mm.PlasticState(st0))
                    assert np.array_equal(a, s1) and np.array_equal(b, t1) and np.array_equal(c.data, n1.data) and np.array_equal(c.iters, n1.iters)
            except BaseException as e:      # noqa: BLE001
                errs.append(e)

        def hyper_worker():
            try:
                for _ in range(50):
                    a, b, _ = mm.material_response(mm.dSdC(), nh, mm.DeformationGradient(F))
                    assert np.array_equal(a, s2) and np.array_equal(b, t2)
            except BaseException as e:      # noqa: BLE001
                errs.append(e)

        def compact_worker():
            try:
                for _ in range(50):
                    a, T, c = mm.material_response(p, eps, mm.PlasticState(st0), tangent="factored")
                    assert np.array_equal(a, s3) and np.array_equal(T.rows, T3.rows) and T.n_rows == T3.n_rows
            except BaseException as e:      # noqa: BLE001
                errs.append(e)

        th = [threading.Thread(target=f) for f in (plastic_worker, hyper_worker, compact_worker)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errs, errs[0]
        assert np.array_equal(T3.dense(), t1) and np.array_equal(n3.data, n1.data)
    finally:
        cuda_lib.mmh_set_chunk_points(old)


def test_pinned_array_views_own_their_memory(mm, torch):
    """a view of a PinnedArray keeps the page-locked block alive after the PinnedArray object (and the original array) are gone"""
    import gc

    a = mm.PinnedArray((6, 1000)).array           # the PinnedArray temporary dies here
    a[...] = 3.0
    v = a[2:4, 10:20]
    del a
    gc.collect()
    junk = [mm.PinnedArray((6, 1000)).array for _ in range(8)]      # would recycle the block if it had been released
    for j in junk:
        j[...] = -1.0
    assert float(v.sum()) == 60.0
    pa = mm.PinnedArray((4,), np.uint16)
    w = pa.array
    pa.free()
    w[:] = 7
    assert int(w.sum()) == 28


def test_strain_and_state_must_live_in_the_same_place(mm, torch):
    """a device state with a host strain (or the reverse) is refused before any pointer reaches the library (it would be an illegal address)"""
    N = 100
    p = mm.Plastic(**sy.PLASTIC_PARAMS)
    eps_h = np.zeros((6, N))
    eps_d = torch.zeros((6, N), dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError, match="same place"):
        mm.material_response(p, eps_h, mm.initial_material_state(p, N))                    # default state lives on the GPU
    with pytest.raises(ValueError, match="same place"):
        mm.material_response(p, eps_d, mm.initial_material_state(p, N, device="cpu"))
    cm = make_crystal(mm)
    with pytest.raises(ValueError, match="same place"):
        mm.material_response(cm, eps_h, mm.initial_material_state(cm, N), 1.0)
    for bad in (torch.zeros((6, N), dtype=torch.float32, device="cuda"), torch.zeros((N, 6), dtype=torch.float64, device="cuda").T):
        with pytest.raises(ValueError):
            mm.material_response(p, eps_d, mm.initial_material_state(p, N), out={"stress": bad})
    torch.cuda.synchronize()


def test_out_buffers_for_every_model(mm, torch):
    """out= (caller-provided, e.g. pinned, buffers) is honoured by every model path and gives the same bits as the allocating call"""
    N = 3001
    le = mm.LinearElastic(E=200e3, ν=0.3)
    nc, lo, hi, seed = sy.c1_strain()
    eps = sy.host_uniform(N, nc, seed, lo, hi)
    ref = mm.material_response(le, eps)
    o = {"stress": mm.PinnedArray((6, N)).array, "tangent": mm.PinnedArray((36, N)).array}
    got = mm.material_response(le, eps, out=o)
    assert got[0] is o["stress"] and np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    cm = make_crystal(mm)
    nc, lo, hi, seed = sy.c4_strain_increment(0)
    de = sy.host_uniform(N, nc, seed, lo, hi)
    st = mm.initial_material_state(cm, N, device="cpu")
    ref = mm.material_response(cm, de, st, 1.0)
    o = {"stress": np.empty((6, N)), "tangent": np.empty((36, N)), "state": np.empty((42, N)), "flags": np.empty(N, np.uint16), "iters": np.empty(N, np.uint16)}
    got = mm.material_response(cm, de, st, 1.0, out=o)
    assert got[2].data is o["state"] and all(np.array_equal(a, b) for a, b in ((got[0], ref[0]), (got[1], ref[1]), (got[2].data, ref[2].data), (got[2].flags, ref[2].flags), (got[2].iters, ref[2].iters)))
    x = sy.XU_PARAMS
    xm = mm.XuNeedleman(σₘₐₓ=x["sigma_max"], τₘₐₓ=x["tau_max"], Φₙ=mm.xu_needleman_Φₙ(x["sigma_max"], x["delta_n"]),
                        Φₜ=mm.xu_needleman_Φₜ(x["tau_max"], x["delta_t"]), Δₙˢ=x["Delta_n_star"])
    nc, lo, hi, seed = sy.c5_jump()
    D = mm.synth_uniform(N, nc, seed, lo, hi)
    ref = mm.material_response(xm, D)
    o = {"stress": torch.empty_like(D), "tangent": torch.empty((9, N), dtype=torch.float64, device="cuda")}
    got = mm.material_response(xm, D, out=o)
    assert got[0] is o["stress"] and torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])
    nh = mm.NeoHook(λ=1.0, μ=1.0)
    nc, lo, hi, seed = sy.c3_defgrad()
    F = mm.synth_uniform(N, nc, seed, lo, hi)
    ref = mm.material_response(mm.dPdF(), nh, mm.DeformationGradient(F))
    o = {"stress": torch.empty((9, N), dtype=torch.float64, device="cuda"), "tangent": torch.empty((81, N), dtype=torch.float64, device="cuda")}
    got = mm.material_response(mm.dPdF(), nh, mm.DeformationGradient(F), out=o)
    assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]) and got[1] is o["tangent"]
    with pytest.raises(ValueError):
        mm.material_response(xm, D, out={"stress": torch.empty((3, N + 1), dtype=torch.float64, device="cuda")})
