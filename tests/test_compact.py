"""Compact return of the J2 update (include/mmb200.h: mm_plastic_compact / mmh_plastic_compact / mmh_plastic_expand) — the reference's
own return shape, src/Plastic.jl:134-135: elastic points share m.Eᵉ and keep their state, yielded points own a row.
CPU: the host expander against the kernels' host twin (bit-exact) and against the oracle.  GPU: every row kind x layout through
the chunk pipeline, expanded, must be BIT-IDENTICAL to the dense entry point."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
import materialmodels_jl_b200._abi as abi
from materialmodels_jl_b200 import synth as sy

KINDS = [abi.MM_ROWS_TANGENT_STATE, abi.MM_ROWS_FACTORED_STATE, abi.MM_ROWS_TANGENT, abi.MM_ROWS_FACTORED]


def _twin_step(N, step=1):
    """steps 0..step on the host twin: returns eps, state_in, sig, dense tan, parts(13,N), state_out, flags, iters of the last one"""
    tw = H.hosttwin()
    p = H.plastic_c()
    st = np.zeros((14, N))
    for k in range(step + 1):
        nc, lo, hi, seed = sy.c2_strain(k)
        eps = sy.host_uniform(N, nc, seed, lo, hi)
        sig, tan, so = np.zeros((6, N)), np.zeros((36, N)), np.zeros((14, N))
        fl, it = np.zeros(N, np.uint8), np.zeros(N, np.uint16)
        tw.twin_plastic(C.byref(p), H.P(eps), H.P(st), H.P(sig), H.P(tan), H.P(so), H.P(fl), H.P(it), C.c_double(1e-8), 1000, C.c_int64(N))
        if k == step:
            sig2, parts, so2 = np.zeros((6, N)), np.zeros((13, N)), np.zeros((14, N))
            fl2, it2 = np.zeros(N, np.uint8), np.zeros(N, np.uint16)
            tw.twin_plastic_parts(C.byref(p), H.P(eps), H.P(st), H.P(sig2), H.P(parts), H.P(so2), H.P(fl2), H.P(it2), C.c_double(1e-8), 1000, C.c_int64(N))
            assert np.array_equal(sig, sig2) and np.array_equal(so, so2) and np.array_equal(fl, fl2) and np.array_equal(it, it2)
            return p, eps, st, sig, tan, parts, so, fl, it
        st = so


def _rows_from(kind, tan, parts, so, fl):
    y = (fl & 1).astype(bool)
    t = parts[:, y].T if kind in (abi.MM_ROWS_FACTORED_STATE, abi.MM_ROWS_FACTORED) else tan[:, y].T
    if kind in (abi.MM_ROWS_TANGENT_STATE, abi.MM_ROWS_FACTORED_STATE):
        t = np.concatenate([t, so[:, y].T], axis=1)
    return np.ascontiguousarray(t), int(y.sum())


@pytest.mark.parametrize("lay", [0, 1])
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("N", [1, 255, 4097, 50000])
def test_expander_rebuilds_the_dense_arrays_bit_exactly(mm, N, kind, lay):
    """rows cut out of the host twin's result -> mmh_plastic_expand -> the twin's dense tangent and state, bit for bit (the factored
    rows go through the same fma sequence as the device kernels: plastic_tangent_entry_lg)"""
    import materialmodels_jl_b200._lib as L

    lib = L.load()
    p, eps, st, sig, tan, parts, so, fl, it = _twin_step(N)
    rows, n = _rows_from(kind, tan, parts, so, fl)
    assert rows.shape == (n, lib.mm_plastic_row_width(kind))
    with_state = kind in (abi.MM_ROWS_TANGENT_STATE, abi.MM_ROWS_FACTORED_STATE)
    lay_of = (lambda a: a) if lay == 0 else (lambda a: np.ascontiguousarray(a.T))
    sin = lay_of(st)
    T = np.full(lay_of(tan).shape, np.nan)
    S = np.full(sin.shape, np.nan) if with_state else None
    rc = lib.mmh_plastic_expand(C.byref(p), kind, H.P(fl), H.P(rows), n, H.P(sin), H.P(T), H.P(S), N, lay)
    assert rc == 0, L.last_error()
    assert np.array_equal(T, lay_of(tan))
    if with_state:
        assert np.array_equal(S, lay_of(so))
        inplace = sin.copy()                                     # state_out == state_in: only the yielded points are touched
        assert lib.mmh_plastic_expand(C.byref(p), kind, H.P(fl), H.P(rows), n, H.P(inplace), None, H.P(inplace), N, lay) == 0
        assert np.array_equal(inplace, lay_of(so))
    else:
        assert lib.mmh_plastic_expand(C.byref(p), kind, H.P(fl), H.P(rows), n, H.P(sin), None, H.P(sin.copy()), N, lay) == abi.MM_ERR_ARG
    assert lib.mmh_plastic_expand(C.byref(p), kind, H.P(fl), H.P(rows), n + 1, H.P(sin), H.P(T), None, N, lay) == abi.MM_ERR_ARG   # count mismatch is caught
    for threads in (1, 3):                                       # the split over host threads does not change the result
        old = lib.mmh_set_host_threads(threads)
        T2 = np.empty_like(T)
        assert lib.mmh_plastic_expand(C.byref(p), kind, H.P(fl), H.P(rows), n, H.P(sin), H.P(T2), None, N, lay) == 0
        lib.mmh_set_host_threads(0)
        assert np.array_equal(T2, T) and old >= 1


def test_expanded_factored_tangent_matches_the_oracle(mm, oracle):
    """end of the chain on the CPU box: oracle (dense 14x14 Jacobian, LU) vs the 13 numbers of a factored row"""
    import materialmodels_jl_b200._lib as L

    N = 20000
    p, eps, st, sig, tan, parts, so, fl, it = _twin_step(N)
    r = oracle.plastic(p, eps, st)
    assert np.array_equal(fl, r["flags"]) and np.array_equal(it, r["iters"])
    rows, n = _rows_from(abi.MM_ROWS_FACTORED, tan, parts, so, fl)
    T = np.empty((36, N))
    assert L.load().mmh_plastic_expand(C.byref(p), abi.MM_ROWS_FACTORED, H.P(fl), H.P(rows), n, None, H.P(T), None, N, 0) == 0
    assert H.rel_err(T, r["tan"]).max() < H.RTOL
    el = fl == 0
    assert np.array_equal(T[:, el], np.repeat(mm.Plastic(**sy.PLASTIC_PARAMS).Eᵉ[:, None], int(el.sum()), axis=1))   # Plastic.jl:135: the stored Eᵉ


def test_compact_tangent_getitem_matches_dense(mm):
    N = 3000
    p, eps, st, sig, tan, parts, so, fl, it = _twin_step(N)
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    from materialmodels_jl_b200.api import CompactTangent

    for kind, name in ((abi.MM_ROWS_FACTORED_STATE, "factored"), (abi.MM_ROWS_TANGENT_STATE, "compact")):
        rows, n = _rows_from(kind, tan, parts, so, fl)
        ct = CompactTangent(m, name, kind, fl, rows, n, "soa")
        assert np.array_equal(ct.dense(), tan)
        for i in list(range(0, N, 97)):
            assert np.abs(ct[i] - tan[:, i]).max() <= 1e-12 * np.abs(tan[:, i]).max()


# ------------------------------------------------------------------------------------------------------------------- GPU
def _dense_reference(mm, m, eps_h, st_h, lay):
    import torch

    layout = "soa" if lay == 0 else "aos"
    e = torch.from_numpy(eps_h).cuda()
    s = mm.PlasticState(torch.from_numpy(st_h).cuda(), layout)
    sig, tan, new = mm.material_response(m, e, s)
    return sig.cpu().numpy(), tan.cpu().numpy(), new.data.cpu().numpy(), new.flags.cpu().numpy(), new.iters.cpu().numpy().view(np.uint16)


def _two_step_inputs(mm, m, N, lay):
    """host arrays (eps of step B, state after step A) in the requested layout"""
    import torch

    nc, lo, hi, seed = sy.c2_strain(0)
    eA = mm.synth_uniform(N, nc, seed, lo, hi)
    _, _, stA = mm.material_response(m, eA, mm.initial_material_state(m, N), tangent=False)
    nc, lo, hi, seed = sy.c2_strain(1)
    eB = sy.host_uniform(N, nc, seed, lo, hi)
    sA = stA.data.cpu().numpy()
    torch.cuda.synchronize()
    if lay == 1:
        return np.ascontiguousarray(eB.T), np.ascontiguousarray(sA.T)
    return eB, sA


@pytest.mark.gpu
@pytest.mark.parametrize("lay", [0, 1])
@pytest.mark.parametrize("kind", [abi.MM_ROWS_TANGENT_STATE, abi.MM_ROWS_FACTORED_STATE])
@pytest.mark.parametrize("N,chunk", [(1, 1 << 20), (777, 100), (40000, 4096), (300001, 1 << 16)])
def test_compact_pipeline_expands_to_the_dense_result(mm, cuda_lib, N, chunk, kind, lay):
    """mmh_plastic_compact (chunked, rows issued LAG chunks late) + mmh_plastic_expand == mm_plastic, bit for bit: σ, flags, iters, tangent, state"""
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    eps, sin = _two_step_inputs(mm, m, N, lay)
    sig_d, tan_d, st_d, fl_d, it_d = _dense_reference(mm, m, eps, sin, lay)
    lib = cuda_lib
    W = lib.mm_plastic_row_width(kind)
    old = lib.mmh_set_chunk_points(chunk)
    try:
        sig, fl, it = np.full(eps.shape, np.nan), np.zeros(N, np.uint8), np.zeros(N, np.uint16)
        rows = np.full((N, W), np.nan)
        nrows, nfail = C.c_int64(-1), np.zeros(1, np.int32)
        rc = lib.mmh_plastic_compact(C.byref(m._c), H.P(eps), H.P(sin), H.P(sig), H.P(fl), H.P(it), kind, H.P(rows), N, C.byref(nrows), None, H.P(nfail), None, N, lay)
        assert rc == 0, mm._lib.last_error()
        n = nrows.value
        assert n == int((fl_d & 1).sum()) and nfail[0] == 0
        assert np.array_equal(sig, sig_d) and np.array_equal(fl, fl_d) and np.array_equal(it, it_d)
        assert not np.isnan(rows[:n]).any() and np.isnan(rows[n:]).all()           # exactly n rows written
        T, S = np.empty_like(tan_d), np.empty_like(st_d)
        assert lib.mmh_plastic_expand(C.byref(m._c), kind, H.P(fl), H.P(rows), n, H.P(sin), H.P(T), H.P(S), N, lay) == 0
        assert np.array_equal(T, tan_d)
        assert np.array_equal(S, st_d)
        # dense state_out straight from the pipeline (may alias state_in)
        S2 = sin.copy()
        rc = lib.mmh_plastic_compact(C.byref(m._c), H.P(eps), H.P(S2), H.P(sig), H.P(fl), H.P(it), kind, H.P(rows), N, C.byref(nrows), H.P(S2), H.P(nfail), None, N, lay)
        assert rc == 0 and np.array_equal(S2, st_d)
        # too small a row buffer: MM_ERR_CAPACITY, the count says what is needed, the per-point outputs are complete
        if n > 1:
            small = np.full((n - 1, W), np.nan)
            sig[:] = np.nan
            rc = lib.mmh_plastic_compact(C.byref(m._c), H.P(eps), H.P(sin), H.P(sig), H.P(fl), H.P(it), kind, H.P(small), n - 1, C.byref(nrows), None, H.P(nfail), None, N, lay)
            assert rc == abi.MM_ERR_CAPACITY and np.array_equal(sig, sig_d)
    finally:
        lib.mmh_set_chunk_points(old)


@pytest.mark.gpu
@pytest.mark.parametrize("lay", [0, 1])
@pytest.mark.parametrize("kind", [abi.MM_ROWS_TANGENT, abi.MM_ROWS_FACTORED])
def test_compact_resident_state(mm, cuda_lib, kind, lay):
    """state kept on the device and updated in place; only ε up, σ / flags / iters / tangent rows down"""
    import torch

    N = 150001
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    eps, sin = _two_step_inputs(mm, m, N, lay)
    sig_d, tan_d, st_d, fl_d, it_d = _dense_reference(mm, m, eps, sin, lay)
    lib = cuda_lib
    W = lib.mm_plastic_row_width(kind)
    old = lib.mmh_set_chunk_points(1 << 15)
    try:
        d_state = torch.from_numpy(sin).cuda()
        sig, fl, it = np.empty_like(eps), np.zeros(N, np.uint8), np.zeros(N, np.uint16)
        rows = np.empty((N, W))
        nrows, nfail = C.c_int64(-1), np.zeros(1, np.int32)
        torch.cuda.synchronize()
        rc = lib.mmh_plastic_compact_resident(C.byref(m._c), H.P(eps), C.c_void_p(d_state.data_ptr()), H.P(sig), H.P(fl), H.P(it), kind, H.P(rows), N, C.byref(nrows),
                                              H.P(nfail), None, N, lay)
        assert rc == 0, mm._lib.last_error()
        n = nrows.value
        assert np.array_equal(sig, sig_d) and np.array_equal(fl, fl_d) and np.array_equal(it, it_d)
        assert np.array_equal(d_state.cpu().numpy(), st_d)
        T = np.empty_like(tan_d)
        assert lib.mmh_plastic_expand(C.byref(m._c), kind, H.P(fl), H.P(rows), n, None, H.P(T), None, N, lay) == 0
        assert np.array_equal(T, tan_d)
        assert lib.mmh_plastic_compact_resident(C.byref(m._c), H.P(eps), C.c_void_p(d_state.data_ptr()), H.P(sig), H.P(fl), H.P(it), abi.MM_ROWS_FACTORED_STATE, H.P(rows), N,
                                                C.byref(nrows), H.P(nfail), None, N, lay) == abi.MM_ERR_ARG
    finally:
        lib.mmh_set_chunk_points(old)


@pytest.mark.gpu
@pytest.mark.parametrize("lay", [0, 1])
def test_device_compact_entry_point(mm, cuda_lib, lay):
    """mm_plastic_compact on device pointers (the building block): rows + count on the device, dense state_out optional"""
    import torch

    N = 100003
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    eps, sin = _two_step_inputs(mm, m, N, lay)
    sig_d, tan_d, st_d, fl_d, it_d = _dense_reference(mm, m, eps, sin, lay)
    lib = cuda_lib
    e, s = torch.from_numpy(eps).cuda(), torch.from_numpy(sin).cuda()
    P = lambda t: C.c_void_p(t.data_ptr())
    for kind in KINDS:
        W = lib.mm_plastic_row_width(kind)
        for want_state in (False, True):
            ws = torch.empty(lib.mm_plastic_compact_workspace(N, int(want_state)), dtype=torch.uint8, device="cuda")
            sig = torch.empty_like(e)
            fl, it = torch.zeros(N, dtype=torch.uint8, device="cuda"), torch.zeros(N, dtype=torch.int16, device="cuda")
            rows = torch.full((N, W), float("nan"), dtype=torch.float64, device="cuda")
            nrows, nfail = torch.zeros(1, dtype=torch.int64, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
            so = torch.empty_like(s) if want_state else None
            rc = lib.mm_plastic_compact(C.byref(m._c), P(e), P(s), P(sig), P(fl), P(it), kind, P(rows), N, P(nrows), P(so) if want_state else None, P(nfail), None,
                                        P(ws), ws.numel(), N, lay, None)
            assert rc == 0, mm._lib.last_error()
            torch.cuda.synchronize()
            n = int(nrows.item())
            assert n == int((fl_d & 1).sum())
            assert np.array_equal(sig.cpu().numpy(), sig_d) and np.array_equal(fl.cpu().numpy(), fl_d)
            if want_state:
                assert np.array_equal(so.cpu().numpy(), st_d)
            T = np.empty_like(tan_d)
            with_state = kind in (abi.MM_ROWS_TANGENT_STATE, abi.MM_ROWS_FACTORED_STATE)
            S = np.empty_like(st_d) if with_state else None
            r = np.ascontiguousarray(rows[:n].cpu().numpy())
            assert not np.isnan(r).any()
            assert lib.mmh_plastic_expand(C.byref(m._c), kind, H.P(fl_d), H.P(r), n, H.P(sin), H.P(T), H.P(S), N, lay) == 0
            assert np.array_equal(T, tan_d)
            if with_state:
                assert np.array_equal(S, st_d)
    # workspace too small is an argument error, not a crash
    ws = torch.empty(256, dtype=torch.uint8, device="cuda")
    assert lib.mm_plastic_compact(C.byref(m._c), P(e), P(s), P(sig), P(fl), P(it), 0, P(rows), N, P(nrows), None, P(nfail), None, P(ws), 256, N, lay, None) == abi.MM_ERR_ARG


@pytest.mark.gpu
def test_python_mirror_compact_forms(mm, cuda_lib):
    """material_response(m, ε, state, tangent="compact" | "factored") with numpy strains; host state and resident (device) state"""
    import torch

    N = 60000
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    eps, sin = _two_step_inputs(mm, m, N, 0)
    sig_d, tan_d, st_d, fl_d, it_d = _dense_reference(mm, m, eps, sin, 0)
    for kind in ("compact", "factored"):
        σ, T, new = mm.material_response(m, eps, mm.PlasticState(sin.copy()), tangent=kind)
        assert isinstance(T, mm.CompactTangent) and T.n_rows == int((fl_d & 1).sum())
        assert np.array_equal(σ, sig_d) and np.array_equal(T.dense(), tan_d) and np.array_equal(new.data, st_d)
        assert np.array_equal(new.iters, it_d) and np.array_equal(T.elastic, m.Eᵉ)
        for i in (0, 1, 17, N - 1):
            assert np.abs(T[i] - tan_d[:, i]).max() <= 1e-12 * np.abs(tan_d[:, i]).max()
        # in-place commit into the caller's state array
        mine = sin.copy()
        _, _, new = mm.material_response(m, eps, mm.PlasticState(mine), tangent=kind, out={"state": mine})
        assert np.array_equal(mine, st_d)
        # resident mode: state on the device, updated in place
        dstate = mm.PlasticState(torch.from_numpy(sin).cuda())
        σ, T, new = mm.material_response(m, eps, dstate, tangent=kind)
        assert np.array_equal(σ, sig_d) and np.array_equal(T.dense(), tan_d)
        assert np.array_equal(new.data.cpu().numpy(), st_d) and new.data.data_ptr() == dstate.data.data_ptr()
    with pytest.raises(BufferError):
        mm.material_response(m, eps, mm.PlasticState(sin.copy()), tangent="factored", out={"rows": np.empty((5, 27))})
    with pytest.raises(ValueError):
        mm.material_response(m, eps, mm.PlasticState(sin.copy()), tangent="sparse")


@pytest.mark.gpu
@pytest.mark.parametrize("lay", [0, 1])
@pytest.mark.parametrize("N,chunk", [(1, 1 << 20), (777, 100), (40000, 4096), (300001, 1 << 16), (2000003, 1 << 18)])
def test_dense_host_call_through_factored_rows_is_the_dense_result(mm, cuda_lib, N, chunk, lay):
    """mmh_plastic with a dense tangent (the default host-buffer J2 call): factored rows over the link + expansion on a helper thread while
    the next chunks are in flight == plain dense transfer == mm_plastic, bit for bit; without flags / iters arrays, with the state updated
    in place, for an all-elastic and an all-yielded batch (every chunk with 0 rows / n rows)."""
    m = mm.Plastic(**sy.PLASTIC_PARAMS)
    eps, sin = _two_step_inputs(mm, m, N, lay)
    sig_d, tan_d, st_d, fl_d, it_d = _dense_reference(mm, m, eps, sin, lay)
    lib = cuda_lib
    old = lib.mmh_set_chunk_points(chunk)

    def call(via, e, s_in, with_flags=True, inplace=False):
        prev = lib.mm_set_tuning(b"mmh_plastic_via_factored", via)
        try:
            sig, T = np.full(e.shape, np.nan), np.full(tan_d.shape, np.nan)
            S = s_in.copy() if inplace else np.full(s_in.shape, np.nan)
            fl, it, nf = np.zeros(N, np.uint8), np.zeros(N, np.uint16), np.zeros(1, np.int32)
            rc = lib.mmh_plastic(C.byref(m._c), H.P(e), H.P(S if inplace else s_in), H.P(sig), H.P(T), H.P(S), H.P(fl) if with_flags else None,
                                 H.P(it) if with_flags else None, H.P(nf), None, N, lay)
            assert rc == 0, mm._lib.last_error()
            return sig, T, S, fl, it, int(nf[0])
        finally:
            lib.mm_set_tuning(b"mmh_plastic_via_factored", prev)

    try:
        for via in (2, 1, 0):          # 2: the factored route whatever the batch size (1 leaves small batches to the plain transfer)
            sig, T, S, fl, it, nf = call(via, eps, sin)
            assert nf == 0 and np.array_equal(sig, sig_d) and np.array_equal(T, tan_d) and np.array_equal(S, st_d) and np.array_equal(fl, fl_d) and np.array_equal(it, it_d)
        sig, T, S, _, _, _ = call(2, eps, sin, with_flags=False)
        assert np.array_equal(sig, sig_d) and np.array_equal(T, tan_d) and np.array_equal(S, st_d)
        sig, T, S, _, _, _ = call(2, eps, sin, inplace=True)
        assert np.array_equal(T, tan_d) and np.array_equal(S, st_d)
        for scale in (1e-3, 50.0):                                  # nobody yields / everybody yields
            e2 = eps * scale
            a, b = call(2, e2, sin), call(0, e2, sin)
            assert (a[3] & 1).mean() in (0.0, 1.0) or N < 10
            assert all(np.array_equal(x, y) for x, y in zip(a[:5], b[:5])) and a[5] == b[5]
    finally:
        lib.mmh_set_chunk_points(old)
