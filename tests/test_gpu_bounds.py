"""Out-of-bounds guard: every compute entry point of the C ABI, both layouts, ragged and tiny N, aligned and
mis-aligned (8- but not 16-byte) array bases — outputs live inside a larger sentinel-filled buffer and the test checks
that (i) not one byte outside the output range was written and (ii) every output element was written.
compute-sanitizer is not available on the GPU pool, so this is the memory-safety net for the ragged-tile / TMA paths
(AoS full tiles go through cp.async.bulk, the ragged last tile and mis-aligned bases through element-wise copies)."""
import ctypes as C

import numpy as np
import pytest

import materialmodels_jl_b200._abi as abi
import materialmodels_jl_b200.synth as sy

pytestmark = pytest.mark.gpu

PAD = 512            # sentinel doubles on each side
SENT = -7.25e300     # never produced by any model on these inputs


class Guarded:
    """an output array of n elements placed `shift` elements into a sentinel-filled torch buffer"""

    def __init__(self, torch, n, dtype, shift):
        self.torch, self.n, self.shift = torch, n, PAD + shift
        self.buf = torch.empty(n + 2 * PAD + 8, dtype=dtype, device="cuda")
        self.sent = SENT if dtype == torch.float64 else 0x5A
        self.buf.fill_(self.sent)
        self.ptr = C.c_void_p(self.buf.data_ptr() + self.shift * self.buf.element_size())

    def check(self, name, all_written=True):
        b = self.buf
        assert bool((b[: self.shift] == self.sent).all()) and bool((b[self.shift + self.n:] == self.sent).all()), "%s: write outside the output range" % name
        if all_written and self.buf.dtype == self.torch.float64:
            assert not bool((b[self.shift: self.shift + self.n] == self.sent).any()), "%s: output element left unwritten" % name
        return b[self.shift: self.shift + self.n]


def inp(torch, a, shift):
    """input array copied `shift` doubles into its own buffer (to produce 8-byte-but-not-16-byte aligned bases)"""
    t = torch.empty(a.size + 8, dtype=torch.float64, device="cuda")
    t[shift: shift + a.size] = torch.from_numpy(np.ascontiguousarray(a).ravel()).cuda()
    return t, C.c_void_p(t.data_ptr() + 8 * shift)


def lay_of(a_soa, lay):
    return a_soa if lay == 0 else np.ascontiguousarray(a_soa.T)


@pytest.mark.parametrize("shift", [0, 1])
@pytest.mark.parametrize("lay", [0, 1])
@pytest.mark.parametrize("N", [1, 31, 129, 1000, 4099])
def test_no_entry_point_writes_out_of_bounds(mm, cuda_lib, N, lay, shift):
    import torch

    lib = cuda_lib
    f64, u8, i16 = torch.float64, torch.uint8, torch.int16
    G = lambda n, dt=f64: Guarded(torch, n, dt, shift)
    keep = []

    def I(a):
        t, p = inp(torch, lay_of(a, lay), shift)
        keep.append(t)
        return p

    nfail = torch.zeros(1, dtype=torch.int32, device="cuda")
    pn = C.c_void_p(nfail.data_ptr())
    # LinearElastic
    le = abi.MMLinearElastic(200e3, 0.3)
    eps = sy.host_uniform(N, 6, 3, [-2e-3] * 6, [2e-3] * 6)
    s, t = G(6 * N), G(36 * N)
    assert lib.mm_linear_elastic(C.byref(le), I(eps), s.ptr, t.ptr, N, lay, None) == 0
    s.check("mm_linear_elastic σ"), t.check("mm_linear_elastic tangent")
    psi = G(N)
    assert lib.mm_linear_elastic_energy(C.byref(le), I(eps), psi.ptr, N, lay, None) == 0
    psi.check("mm_linear_elastic_energy")
    # Plastic (virgin state, ~half the points yield)
    pp = abi.MMPlastic(*[sy.PLASTIC_PARAMS[k] for k in ("E", "nu", "sigma_y", "H", "r", "kappa_inf", "alpha_inf")])
    nc, lo, hi, seed = sy.c2_strain(0)
    e2 = sy.host_uniform(N, nc, seed, lo, hi)
    s, t, so, fl, it = G(6 * N), G(36 * N), G(14 * N), G(N, u8), G(N, i16)
    assert lib.mm_plastic(C.byref(pp), I(e2), I(np.zeros((14, N))), s.ptr, t.ptr, so.ptr, fl.ptr, it.ptr, pn, None, N, lay, None) == 0
    for g, n in ((s, "σ"), (t, "tangent"), (so, "state"), (fl, "flags"), (it, "iters")):
        g.check("mm_plastic " + n)
    # compact return: rows sized EXACTLY for the yielded points (one more row would hit the sentinel), guarded count and workspace
    n_yield = int((fl.check("flags").cpu().numpy() & 1).sum())
    for kind in (abi.MM_ROWS_TANGENT_STATE, abi.MM_ROWS_FACTORED_STATE, abi.MM_ROWS_TANGENT, abi.MM_ROWS_FACTORED):
        W = lib.mm_plastic_row_width(kind)
        wsz = lib.mm_plastic_compact_workspace(N, 1)
        s, so, fl2, it2, rows, ws = G(6 * N), G(14 * N), G(N, u8), G(N, i16), G(max(n_yield, 1) * W), G((wsz + 7) // 8)
        cnt = torch.full((3,), -1, dtype=torch.int64, device="cuda")
        args = (C.byref(pp), I(e2), I(np.zeros((14, N))), s.ptr, fl2.ptr, it2.ptr, kind, rows.ptr, n_yield, C.c_void_p(cnt.data_ptr() + 8), so.ptr, pn, None, ws.ptr)
        assert lib.mm_plastic_compact(*args, wsz - 256, N, lay, None) == abi.MM_ERR_ARG          # a workspace declared too small is refused
        assert lib.mm_plastic_compact(*args, wsz, N, lay, None) == 0
        torch.cuda.synchronize()
        assert cnt.tolist() == [-1, n_yield, -1]
        for g, n in ((s, "σ"), (so, "state"), (fl2, "flags"), (it2, "iters")):
            g.check("mm_plastic_compact " + n)
        ws.check("mm_plastic_compact workspace", all_written=False)      # nothing outside the declared workspace range
        if n_yield:
            rows.check("mm_plastic_compact rows")
    # wrapped Plastic: PlaneStress (nc 3) and UniaxialStress (nc 1)
    for kind, ncr, amp in ((abi.MM_DIM_PLANE_STRESS, 3, 1.2e-3), (abi.MM_DIM_UNIAXIAL_STRESS, 1, 2.5e-3), (abi.MM_DIM_PLANE_STRAIN, 3, 1e-3)):
        er = sy.host_uniform(N, ncr, 31, [-amp] * ncr, [amp] * ncr)
        s, t, so, fl, it, oi = G(ncr * N), G(ncr * ncr * N), G(14 * N), G(N, u8), G(N, i16), G(N, u8)
        assert lib.mm_wrapped_plastic(kind, C.byref(pp), I(er), I(np.zeros((14, N))), s.ptr, t.ptr, so.ptr, fl.ptr, it.ptr, oi.ptr, pn, None, None, N, lay, None) == 0
        for g, n in ((s, "σ"), (t, "tangent"), (so, "state"), (fl, "flags"), (it, "iters"), (oi, "outer iters")):
            g.check("mm_wrapped_plastic[%d] %s" % (kind, n))
        s, t, oi = G(ncr * N), G(ncr * ncr * N), G(N, u8)
        assert lib.mm_wrapped_linear_elastic(kind, C.byref(le), I(er), s.ptr, t.ptr, oi.ptr, pn, None, N, lay, None) == 0
        s.check("mm_wrapped_linear_elastic σ"), t.check("mm_wrapped_linear_elastic tangent"), oi.check("mm_wrapped_linear_elastic outer iters")
    # hyperelastic, all output kinds
    hp = abi.MMHyper(**sy.YEOH_PARAMS)
    nc, lo, hi, seed = sy.c3_defgrad()
    F = sy.host_uniform(N, nc, seed, lo, hi)
    for tk, ns, nt in ((abi.MM_TANGENT_DSDC, 6, 36), (abi.MM_TANGENT_DSDE, 6, 36), (abi.MM_TANGENT_DPDF, 9, 81)):
        s, t = G(ns * N), G(nt * N)
        assert lib.mm_hyper_ex(abi.MM_YEOH, C.byref(hp), I(F), abi.MM_STRAIN_F, tk, s.ptr, t.ptr, N, lay, None) == 0
        s.check("mm_hyper_ex stress"), t.check("mm_hyper_ex tangent")
    psi = G(N)
    assert lib.mm_hyper_energy(abi.MM_YEOH, C.byref(hp), I(F), abi.MM_STRAIN_F, psi.ptr, N, lay, None) == 0
    psi.check("mm_hyper_energy")
    # XuNeedleman 2D / 3D
    xp = abi.MMXuNeedleman()
    x = sy.XU_PARAMS
    assert lib.mm_xu_needleman_setup(x["sigma_max"], x["tau_max"], x["sigma_max"] * np.e * x["delta_n"], x["tau_max"] * np.sqrt(0.5 * np.e) * x["delta_t"],
                                     x["Delta_n_star"], C.byref(xp)) == 0
    for dim in (2, 3):
        D = sy.host_uniform(N, dim, 9, [-0.004] * (dim - 1) + [-0.001], [0.004] * (dim - 1) + [0.008])
        T, H = G(dim * N), G(dim * dim * N)
        assert lib.mm_xu_needleman(C.byref(xp), I(D), T.ptr, H.ptr, dim, N, lay, None) == 0
        T.check("mm_xu_needleman T"), H.check("mm_xu_needleman dTdΔ")
        if lay == 0:                                   # the copy-engine tile kernel for SoA arrays (A/B option)
            T, H = G(dim * N), G(dim * dim * N)
            old = lib.mm_set_tuning(b"xu_soa_tiled", 1)
            try:
                assert lib.mm_xu_needleman(C.byref(xp), I(D), T.ptr, H.ptr, dim, N, lay, None) == 0
            finally:
                lib.mm_set_tuning(b"xu_soa_tiled", old)
            T.check("mm_xu_needleman T (tiles)"), H.check("mm_xu_needleman dTdΔ (tiles)")
    # TransverselyIsotropic
    tp = abi.MMTransverselyIsotropic()
    assert lib.mm_transversely_isotropic_setup(100e3, 10e3, 5e3, 0.4, 0.3, C.byref(tp)) == 0
    a3 = sy.host_uniform(N, 3, 100, [0.1] * 3, [1.0] * 3)
    a3 /= np.linalg.norm(a3, axis=0)
    s, t = G(6 * N), G(36 * N)
    assert lib.mm_transversely_isotropic(C.byref(tp), I(eps), I(a3), s.ptr, t.ptr, N, lay, None) == 0
    s.check("mm_transversely_isotropic σ"), t.check("mm_transversely_isotropic E")
    # CrystalViscoPlastic: structured two-kernel path and (q != 0) cross-hardening path
    for q in (0.0, 0.3):
        prm = dict(sy.CRYSTAL_PARAMS)
        prm["q"] = q
        cm = mm.CrystalViscoPlastic(slipsystems=mm.slipsystems(mm.FCC(), sy.CRYSTAL_RODRIGUES), **prm)
        nc, lo, hi, seed = sy.c4_strain_increment(0)
        de = sy.host_uniform(N, nc, seed, lo, hi)
        s, t, so, ac, it = G(6 * N), G(36 * N), G(42 * N), G(N, i16), G(N, i16)
        assert lib.mm_crystal(C.byref(cm._c), I(de), 1.0, I(np.zeros((42, N))), s.ptr, t.ptr, so.ptr, ac.ptr, it.ptr, pn, None, N, lay, None) == 0
        for g, n in ((s, "σ"), (t, "tangent"), (so, "state"), (it, "iters"), (ac, "active")):
            g.check("mm_crystal(q=%g) %s" % (q, n))
    # wrapped crystal (general inner path, Jacobian in the shared-memory scratch column)
    for kind, ncr, amp in ((abi.MM_DIM_PLANE_STRESS, 3, 4e-3), (abi.MM_DIM_UNIAXIAL_STRESS, 1, 6e-3)):
        er = sy.host_uniform(N, ncr, 33, [-amp] * ncr, [amp] * ncr)
        s, t, so, ac, it, oi = G(ncr * N), G(ncr * ncr * N), G(42 * N), G(N, i16), G(N, i16), G(N, u8)
        assert lib.mm_wrapped_crystal(kind, C.byref(cm._c), I(er), 1.0, I(np.zeros((42, N))), s.ptr, t.ptr, so.ptr, ac.ptr, it.ptr, oi.ptr, pn, None, None, N,
                                      lay, None) == 0
        for g, n in ((s, "σ"), (t, "tangent"), (so, "state"), (it, "iters"), (ac, "active"), (oi, "outer iters")):
            g.check("mm_wrapped_crystal[%d] %s" % (kind, n))
    # fused cell kernels (SoA only): ncell chosen so that P = ncell*nq is ragged against the 128-thread blocks
    if lay == 0:
        for nn, nq in ((4, 1), (4, 4), (8, 1), (8, 8), (10, 4)):
            ncell = max(1, N // nq)
            P = ncell * nq
            dN = sy.host_uniform(P, 3 * nn, 71, [-1.0] * (3 * nn), [1.0] * (3 * nn))
            dO = sy.host_uniform(P, 1, 72, [0.5], [1.5])
            ue = sy.host_uniform(ncell, 3 * nn, 73, [-4e-4] * (3 * nn), [4e-4] * (3 * nn))
            fe, kk, s, t, so, fl, it = G(3 * nn * ncell), G(9 * nn * nn * ncell), G(6 * P), G(36 * P), G(14 * P), G(P, u8), G(P, i16)
            assert lib.mm_cells_plastic(C.byref(pp), nn, nq, I(dN), I(dO), I(ue), I(np.zeros((14, P))), fe.ptr, kk.ptr, s.ptr, t.ptr, so.ptr, fl.ptr, it.ptr, pn,
                                        None, ncell, None) == 0
            for g, n in ((fe, "fe"), (kk, "ke"), (s, "σ"), (t, "tangent"), (so, "state"), (fl, "flags"), (it, "iters")):
                g.check("mm_cells_plastic(%d,%d) %s" % (nn, nq, n))
            fe, kk, s, t = G(3 * nn * ncell), G(9 * nn * nn * ncell), G(6 * P), G(36 * P)
            a3c = sy.host_uniform(P, 3, 101, [0.1] * 3, [1.0] * 3)
            a3c /= np.linalg.norm(a3c, axis=0)
            assert lib.mm_cells_transversely_isotropic(C.byref(tp), nn, nq, I(dN), I(dO), I(ue), I(a3c), fe.ptr, kk.ptr, s.ptr, t.ptr, ncell, None) == 0
            for g, n in ((fe, "fe"), (kk, "ke"), (s, "σ"), (t, "tangent")):
                g.check("mm_cells_transversely_isotropic(%d,%d) %s" % (nn, nq, n))
            fe, pt, dp = G(3 * nn * ncell), G(9 * P), G(81 * P)
            u2 = sy.host_uniform(ncell, 3 * nn, 75, [-0.03] * (3 * nn), [0.03] * (3 * nn))
            assert lib.mm_cells_hyper(abi.MM_YEOH, C.byref(hp), nn, nq, I(dN), I(dO), I(u2), fe.ptr, pt.ptr, dp.ptr, ncell, None) == 0
            fe.check("mm_cells_hyper fe"), pt.check("mm_cells_hyper Pt"), dp.check("mm_cells_hyper dPdF")
            fe, s = G(3 * nn * ncell), G(6 * P)
            assert lib.mm_cells_linear_elastic(C.byref(le), nn, nq, I(dN), I(dO), I(ue), fe.ptr, None, s.ptr, None, ncell, None) == 0
            fe.check("mm_cells_linear_elastic fe"), s.check("mm_cells_linear_elastic σ")
    torch.cuda.synchronize()
    assert lib.mm_last_error() is not None
