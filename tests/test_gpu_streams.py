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
