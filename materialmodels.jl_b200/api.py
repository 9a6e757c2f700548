"""Batched mirror of the reference's generic API (reference src/MaterialModels.jl:41-81):

    initial_material_state(m, N, ...)                       <- initial_material_state(m)
    get_cache(m)                                            <- get_cache(m)
    material_response(m, strain, state, Δt; cache, options) <- material_response(...)  (one call = N points)
    material_response(dSdC(), m, DeformationGradient(F), state, Δt)   <- src/traits.jl:91-107

Arrays are float64.  `torch` CUDA tensors go straight to the device-pointer entry points (mm_*) on the
current torch stream; numpy arrays (host) go through the pipelined host-buffer entry points (mmh_*).
Layout: "soa" = shape (nc, N) [default, fast path], "aos" = shape (N, nc) (== a Julia
Vector{SymmetricTensor{2,3}} reinterpreted).  Either way the hot path is the CUDA library; if it is
missing, every call raises (no CPU fallback).
"""
import ctypes

import numpy as np

from . import _abi
from ._lib import check, load
from .materials import (AbstractDim, AbstractMaterialState, AbstractTangent, CrystalViscoPlastic, DeformationGradient, GreenLagrange, dsigmadeps,
                        LinearElastic, Plastic, RightCauchyGreen, SmallStrain, StrainMeasure, TransverselyIsotropic, XuNeedleman, _Hyper,
                        dPdF, dSdC, dSdE)

NOT_CONVERGED = "Material model not converged. Could not find material state."   # Plastic.jl:157, CrystalViscoPlastic.jl:92
PLANE_STRESS_NOT_CONVERGED = "Not converged. Could not find plane stress state."    # wrappers.jl:78

_LAYOUTS = {"soa": _abi.MM_LAYOUT_SOA, "aos": _abi.MM_LAYOUT_AOS, 0: 0, 1: 1}


def _torch():
    import torch

    return torch


def _is_host(a):
    return isinstance(a, np.ndarray)


def _shape(nc, N, layout):
    return (nc, N) if _LAYOUTS[layout] == _abi.MM_LAYOUT_SOA else (N, nc)


def _npoints(a, nc, layout):
    shp = tuple(a.shape)
    if _LAYOUTS[layout] == _abi.MM_LAYOUT_SOA:
        if len(shp) != 2 or shp[0] != nc:
            raise ValueError("expected shape (%d, N) for layout 'soa', got %s" % (nc, shp))
        return shp[1]
    if len(shp) != 2 or shp[1] != nc:
        raise ValueError("expected shape (N, %d) for layout 'aos', got %s" % (nc, shp))
    return shp[0]


def _prep(a, like=None):
    """contiguous float64, on host (numpy) or device (torch)"""
    if _is_host(a):
        return np.ascontiguousarray(a, dtype=np.float64)
    torch = _torch()
    if not a.is_cuda:
        raise ValueError("torch tensors must live on a CUDA device (pass numpy arrays for host data)")
    if a.dtype != torch.float64:
        raise ValueError("float64 required")
    return a.contiguous()


def _empty(shape, like, dtype=None):
    if _is_host(like):
        return np.empty(shape, dtype=dtype or np.float64)
    torch = _torch()
    return torch.empty(shape, dtype=dtype or torch.float64, device=like.device)


def _zeros(shape, like, dtype=None):
    if _is_host(like):
        return np.zeros(shape, dtype=dtype or np.float64)
    torch = _torch()
    return torch.zeros(shape, dtype=dtype or torch.float64, device=like.device)


def _ptr(a):
    if a is None:
        return None
    if _is_host(a):
        return a.ctypes.data_as(ctypes.c_void_p)
    return ctypes.c_void_p(a.data_ptr())


def _stream(like):
    """the caller's current CUDA stream on the array's device, as the void* the C ABI takes"""
    torch = _torch()
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)          # same lookup without building a Stream object (~7 us -> < 1 us per call)
    if raw is not None:
        idx = like.device.index
        return ctypes.c_void_p(raw(idx if idx is not None else torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream(like.device).cuda_stream)


class _DeviceGuard:
    def __init__(self, like):
        self.like = like
        self.ctx = None

    def __enter__(self):
        if not _is_host(self.like):
            torch = _torch()
            idx = self.like.device.index
            if idx is not None and idx != torch.cuda.current_device():   # the common single-GPU case needs no device switch
                self.ctx = torch.cuda.device(idx)
                self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


# --------------------------------------------------------------------------------------------
# states
class LinearElasticState(AbstractMaterialState):
    pass


class NeoHookState(AbstractMaterialState):
    pass


class YeohState(AbstractMaterialState):
    pass


class StVenantState(AbstractMaterialState):
    pass


class XuNeedlemanState(AbstractMaterialState):
    pass


class _ArrayState(AbstractMaterialState):
    NC = 0

    def __init__(self, data, layout="soa"):
        self.data = data
        self.layout = layout
        self.N = _npoints(data, self.NC, layout)
        self.n_fail = None      # device/host counter of the call that produced this state (None for initial states)
        self.flags = None
        self.iters = None

    def _comp(self, lo, hi):
        return self.data[lo:hi] if _LAYOUTS[self.layout] == 0 else self.data[:, lo:hi]


class PlasticState(_ArrayState):
    """[εᵖ(6), κ, α(6), μ] per point — reference src/Plastic.jl:41-46."""
    NC = 14
    εᵖ = property(lambda s: s._comp(0, 6))
    κ = property(lambda s: s._comp(6, 7))
    α = property(lambda s: s._comp(7, 13))
    μ = property(lambda s: s._comp(13, 14))


class CrystalViscoPlasticState(_ArrayState):
    """[σ(6), κ(12), α(12), μ(12)] per point — reference CrystalViscoPlastic.jl:52-57."""
    NC = 42
    σ = property(lambda s: s._comp(0, 6))
    κ = property(lambda s: s._comp(6, 18))
    α = property(lambda s: s._comp(18, 30))
    μ = property(lambda s: s._comp(30, 42))


class TransverselyIsotropicState(_ArrayState):
    """a3(3) per point: unit vector normal to the plane of isotropy — reference src/transverselyisotropic.jl:55-57."""
    NC = 3
    a3 = property(lambda s: s.data)


def _transviso_state(a3, N, device, layout):
    """initial_material_state(::TransverselyIsotropic, a3 = (1,0,0)) — reference src/transverselyisotropic.jl:59-64.
    `a3`: one direction (3,) shared by all N points, or a per-point array of shape (3, N) [soa] / (N, 3) [aos]."""
    host = str(device) == "cpu"
    if a3 is None:
        a3 = (1.0, 0.0, 0.0)
    if not _is_host(a3) and not isinstance(a3, (tuple, list)):                # torch tensor: per-point directions
        data = _prep(a3)
        nrm = data.norm(dim=0 if _LAYOUTS[layout] == 0 else 1)
        bad = (nrm - 1.0).abs() > 1.4901161193847656e-08 * _torch().maximum(nrm, _torch().ones_like(nrm))
        if bool(bad.any().item()):
            raise ValueError("The material direction vector is not a unit vector (norm(a3) = %r)" % float(nrm[bad][0].item()))
        return TransverselyIsotropicState(data, layout)
    a = np.asarray(a3, dtype=np.float64)
    if a.ndim == 1:
        if a.shape != (3,):
            raise ValueError("a3 must have 3 components")
        if N is None:
            raise TypeError("initial_material_state(::TransverselyIsotropic, a3, N): number of points required")
        a = np.broadcast_to(a[:, None], (3, N)) if _LAYOUTS[layout] == 0 else np.broadcast_to(a[None, :], (N, 3))
    a = np.ascontiguousarray(a)
    nrm = np.linalg.norm(a, axis=0 if _LAYOUTS[layout] == 0 else 1)
    bad = np.abs(nrm - 1.0) > np.sqrt(np.finfo(np.float64).eps) * np.maximum(nrm, 1.0)     # Julia's isapprox default rtol
    if bad.any():
        raise ValueError("The material direction vector is not a unit vector (norm(a3) = %r)" % float(nrm[bad][0]))
    data = a if host else _torch().from_numpy(a).to(device)
    return TransverselyIsotropicState(data, layout)


def initial_material_state(m, *args, N=None, device="cuda", layout="soa", a3=None):
    """Zero-initialised state for N points (reference: initial_material_state(m), one point):
    initial_material_state(m, N).  device="cpu" gives a numpy-backed state for the host-buffer path.
    TransverselyIsotropic: initial_material_state(m, a3, N) / (m, a3_per_point) / (m, N) [default direction (1,0,0)]."""
    args = list(args)
    if isinstance(m, TransverselyIsotropic):
        if args and not isinstance(args[0], (int, np.integer)):
            a3 = args.pop(0)
        if args:
            N = int(args.pop(0))
        return _transviso_state(a3, N, device, layout)
    if args:
        N = int(args.pop(0))
    if args:
        raise TypeError("initial_material_state: too many positional arguments")
    if isinstance(m, LinearElastic):
        return LinearElasticState()
    if isinstance(m, XuNeedleman):
        return XuNeedlemanState()
    if isinstance(m, _Hyper):
        return {0: NeoHookState, 1: YeohState, 2: StVenantState}[m.MODEL]()
    cls = PlasticState if isinstance(m, Plastic) else CrystalViscoPlasticState if isinstance(m, CrystalViscoPlastic) else None
    if cls is None:
        raise TypeError("unknown material %r" % (m,))
    if N is None:
        raise TypeError("initial_material_state(%s, N): number of points required" % type(m).__name__)
    if str(device) == "cpu":
        data = np.zeros(_shape(cls.NC, N, layout))
    else:
        torch = _torch()
        data = torch.zeros(_shape(cls.NC, N, layout), dtype=torch.float64, device=device)
    return cls(data, layout)


class PlasticCache:
    """Placeholder kept for API compatibility: the reference's cache holds NLsolve scratch
    (src/Plastic.jl:51-63); the GPU path keeps the whole local solve in registers."""


def get_cache(m):
    return PlasticCache() if isinstance(m, Plastic) else None     # reference src/MaterialModels.jl:71-73


def elastic_tangent_3D(E, ν):
    out = np.zeros(36)
    check(load().mm_elastic_tangent_3d(float(E), float(ν), _ptr(out)), "mm_elastic_tangent_3d")
    return out


# --------------------------------------------------------------------------------------------
def _newton_opts(m, options):
    if isinstance(m, Plastic):
        tol, max_iter = 1e-8, 1000        # NLsolve defaults (ftol, iterations)
        if options:
            p = options.get("nlsolve_params", options.get(":nlsolve_params", {})) or {}
            method = p.get("method", "newton")
            if method not in ("newton", ":newton"):
                raise NotImplementedError("only NLsolve method=:newton is implemented on the GPU path")
            tol = float(p.get("ftol", tol))
            max_iter = int(p.get("iterations", max_iter))
    else:
        tol, max_iter = 1e-8, 100         # solve(...; max_iter=100, tol=1e-8), nonlinear_solver.jl:52
        if options:
            opt = options._asdict() if hasattr(options, "_asdict") else dict(options)
            tol = float(opt.get("tol", tol))
            max_iter = int(opt.get("max_iter", max_iter))
    return _abi.MMNewtonOpts(tol, max_iter, 0)


def _finish_fail(nfail, like, options):
    if options and (options.get("defer_check") if isinstance(options, dict) else False):
        return
    n = int(nfail[0]) if _is_host(like) else int(nfail.item())
    if n != 0:
        raise RuntimeError(NOT_CONVERGED)


def _response_linear_elastic(m, ε, state, layout, want_tangent, out=None):
    ε = _prep(ε)
    N = _npoints(ε, 6, layout)
    σ = _out(out, "stress", _shape(6, N, layout), ε)
    tan = _out(out, "tangent", _shape(36, N, layout), ε) if want_tangent else None
    lib = load()
    with _DeviceGuard(ε):
        if _is_host(ε):
            rc = lib.mmh_linear_elastic(ctypes.byref(m._c), _ptr(ε), _ptr(σ), _ptr(tan), N, _LAYOUTS[layout])
        else:
            rc = lib.mm_linear_elastic(ctypes.byref(m._c), _ptr(ε), _ptr(σ), _ptr(tan), N, _LAYOUTS[layout], _stream(ε))
    check(rc, "mm_linear_elastic")
    return σ, (tan if want_tangent else m.Eᵉ), state


def _out(out, key, shape, like, dtype=None):
    """caller-provided output buffer (e.g. pinned host memory) or a fresh allocation.  A provided buffer is written through its raw
    pointer, so everything the native side assumes is checked here: shape, dtype, contiguity, and that it lives where the inputs live."""
    if out is not None and out.get(key) is not None:
        a = out[key]
        if tuple(a.shape) != tuple(shape):
            raise ValueError("out[%r] has shape %s, expected %s" % (key, tuple(a.shape), tuple(shape)))
        if _is_host(like) != _is_host(a):
            raise ValueError("out[%r] must live where the inputs live" % key)
        if _is_host(a):
            want = np.dtype(dtype or np.float64)
            if a.dtype != want:
                raise ValueError("out[%r] has dtype %s, expected %s" % (key, a.dtype, want))
            if not a.flags["C_CONTIGUOUS"] or not a.flags["WRITEABLE"]:
                raise ValueError("out[%r] must be a writeable C-contiguous array" % key)
        else:
            torch = _torch()
            want = dtype or torch.float64
            if a.dtype != want:
                raise ValueError("out[%r] has dtype %s, expected %s" % (key, a.dtype, want))
            if not a.is_contiguous():
                raise ValueError("out[%r] must be contiguous" % key)
            if a.device != like.device:
                raise ValueError("out[%r] is on %s, the inputs are on %s" % (key, a.device, like.device))
        return a
    return _empty(shape, like, dtype)


def _same_place(strain, state_data, what):
    """strain and state must both be host arrays or both live on the same CUDA device (a device pointer handed to mmh_* as a host
    source, or a host pointer handed to a kernel, is an illegal address that poisons the CUDA context)"""
    if _is_host(strain) != _is_host(state_data):
        raise ValueError("%s: strain and state must live in the same place (strain on %s, state on %s); initial_material_state(m, N, device='cpu') "
                         "gives a host state for numpy strains" % (what, "host" if _is_host(strain) else strain.device, "host" if _is_host(state_data) else state_data.device))
    if not _is_host(strain) and strain.device != state_data.device:
        raise ValueError("%s: strain is on %s, state on %s" % (what, strain.device, state_data.device))


class CompactTangent:
    """∂σ∂ε of a batched J2 update in the reference's own terms (src/Plastic.jl:134-135): every elastic point shares the ONE stored
    m.Eᵉ (`elastic`, 36), every yielded point (flags & MM_FLAG_PLASTIC) owns one row of `rows`, in point order.
    kind "compact": the row is the 36-entry tangent; "factored": the row is (gξ, u(6), n(6)) with ∂σ∂ε = Eᵉ − gξ I_dev + u⊗n.
    `tangent[i]` gives the 36 entries of point i; `dense()` materialises the (36, N) / (N, 36) array (bit-identical to the dense path)."""

    def __init__(self, m, kind, row_kind, flags, rows, n_rows, layout, state_in=None):
        self.m, self.kind, self.row_kind, self.flags, self.rows, self.n_rows, self.layout = m, kind, row_kind, flags, rows, n_rows, layout
        self.N = flags.shape[0]
        self.elastic = m.Eᵉ
        self.width = rows.shape[1] if rows.ndim == 2 else 0
        self._state_in = state_in
        self._row_of = None

    @property
    def row_of(self):
        """row index per point, -1 for elastic points"""
        if self._row_of is None:
            y = (self.flags & _abi.MM_FLAG_PLASTIC).astype(bool)
            r = np.cumsum(y, dtype=np.int64) - 1
            r[~y] = -1
            self._row_of = r
        return self._row_of

    def __getitem__(self, i):
        k = int(self.row_of[i])
        if k < 0:
            return self.elastic.copy()
        r = self.rows[k]
        if self.kind == "compact":
            return np.array(r[:36])
        gxi, u, n = r[0], r[1:7], r[7:13]
        idev = _IDEV36
        return (self.elastic - gxi * idev + np.outer(n, u).reshape(36))      # entry I + 6J = u[I] n[J]

    def dense(self, out=None):
        shape = _shape(36, self.N, self.layout)
        tan = np.empty(shape) if out is None else out
        if tuple(tan.shape) != shape or tan.dtype != np.float64 or not tan.flags["C_CONTIGUOUS"]:
            raise ValueError("dense(out=...): expected a C-contiguous float64 array of shape %s" % (shape,))
        check(load().mmh_plastic_expand(ctypes.byref(self.m._c), self.row_kind, _ptr(self.flags), _ptr(self.rows), self.n_rows, None, _ptr(tan), None,
                                        self.N, _LAYOUTS[self.layout]), "mmh_plastic_expand")
        return tan


def _idev36():
    d = np.zeros((6, 6))
    diag = (0, 3, 5)
    for I in range(6):
        for J in range(6):
            if I in diag and J in diag:
                d[J, I] = 2.0 / 3.0 if I == J else -1.0 / 3.0
            elif I == J:
                d[J, I] = 0.5
    return d.reshape(36)


_IDEV36 = _idev36()


class CompactPlasticState(_ArrayState):
    """New state of a compact J2 update: elastic points keep the caller's previous state (the reference returns the same object,
    Plastic.jl:135), yielded points own the trailing 14 entries of their row.  `.data` materialises the dense array on first use;
    `commit(into)` scatters the yielded rows into an existing dense state array (in place when `into` is the previous state's array)."""
    NC = 14

    def __init__(self, m, prev, row_kind, flags, rows, n_rows, layout):
        self.m, self.prev, self.row_kind, self.rows, self.n_rows = m, prev, row_kind, rows, n_rows
        self.layout = layout
        self.N = flags.shape[0]
        self.flags = flags
        self.iters = None
        self.n_fail = None
        self._data = None

    def commit(self, into=None):
        prev = np.ascontiguousarray(self.prev)
        if into is None:
            into = np.empty_like(prev)
        check(load().mmh_plastic_expand(ctypes.byref(self.m._c), self.row_kind, _ptr(self.flags), _ptr(self.rows), self.n_rows, _ptr(prev), None, _ptr(into),
                                        self.N, _LAYOUTS[self.layout]), "mmh_plastic_expand")
        self._data = into
        return into

    @property
    def data(self):
        if self._data is None:
            self.commit()
        return self._data

    @data.setter
    def data(self, v):
        self._data = v


_ROW_KINDS = {("compact", True): _abi.MM_ROWS_TANGENT_STATE, ("factored", True): _abi.MM_ROWS_FACTORED_STATE,
              ("compact", False): _abi.MM_ROWS_TANGENT, ("factored", False): _abi.MM_ROWS_FACTORED}


def _response_plastic_compact(m, ε, state, options, kind, out):
    """material_response(m::Plastic, ε, state; tangent="compact" | "factored") for HOST strains: the compact return of
    include/mmb200.h (mmh_plastic_compact).  state on the host: rows carry [tangent | new state]; state on a CUDA device (resident
    mode): the device state is updated IN PLACE and only ε goes up / σ, flags, iters and tangent rows come down.
    out: stress (6,N), flags (N,) uint8, iters (N,) uint16, rows (capacity, W) float64 [W = mm_plastic_row_width]."""
    layout = state.layout
    ε = _prep(ε)
    if not _is_host(ε):
        raise NotImplementedError("tangent=%r is the host-buffer form of the update (numpy strains); device callers use mm_plastic_compact through the C ABI" % kind)
    N = _npoints(ε, 6, layout)
    if N != state.N:
        raise ValueError("strain has %d points, state has %d" % (N, state.N))
    resident = not _is_host(state.data)
    row_kind = _ROW_KINDS[(kind, not resident)]
    lib = load()
    W = lib.mm_plastic_row_width(row_kind)
    σ = _out(out, "stress", _shape(6, N, layout), ε)
    flags = _out(out, "flags", (N,), ε, np.uint8)
    iters = _out(out, "iters", (N,), ε, np.uint16)
    rows = out.get("rows") if out else None
    if rows is None:
        rows = np.empty((N, W))
    if rows.ndim != 2 or rows.shape[1] != W or rows.dtype != np.float64 or not rows.flags["C_CONTIGUOUS"]:
        raise ValueError("out['rows'] must be a C-contiguous float64 array of shape (capacity, %d)" % W)
    nrows = ctypes.c_int64(0)
    nfail = np.zeros(1, np.int32)
    o = _newton_opts(m, options)
    if resident:
        sdev = state.data
        if not sdev.is_contiguous() or sdev.dtype != _torch().float64:
            raise ValueError("resident state must be a contiguous float64 CUDA tensor")
        with _DeviceGuard(sdev):
            _torch().cuda.current_stream(sdev.device).synchronize()       # the pipeline runs on the library's own streams
            rc = lib.mmh_plastic_compact_resident(ctypes.byref(m._c), _ptr(ε), _ptr(sdev), _ptr(σ), _ptr(flags), _ptr(iters), row_kind, _ptr(rows), rows.shape[0],
                                                  ctypes.byref(nrows), _ptr(nfail), ctypes.byref(o), N, _LAYOUTS[layout])
    else:
        sin = _prep(state.data)
        rc = lib.mmh_plastic_compact(ctypes.byref(m._c), _ptr(ε), _ptr(sin), _ptr(σ), _ptr(flags), _ptr(iters), row_kind, _ptr(rows), rows.shape[0],
                                     ctypes.byref(nrows), None, _ptr(nfail), ctypes.byref(o), N, _LAYOUTS[layout])
    if rc == _abi.MM_ERR_CAPACITY:
        raise BufferError("out['rows'] holds %d rows but %d points yielded" % (rows.shape[0], nrows.value))
    check(rc, "mmh_plastic_compact")
    n = int(nrows.value)
    tan = CompactTangent(m, kind, row_kind, flags, rows[:n], n, layout)
    if resident:
        new = PlasticState(state.data, layout)                 # updated in place on the device
        new.flags = flags
    else:
        new = CompactPlasticState(m, sin, row_kind, flags, rows[:n], n, layout)
        if out and out.get("state") is not None:
            new.commit(_out(out, "state", _shape(14, N, layout), ε))
    new.iters, new.n_fail = iters, nfail
    _finish_fail(nfail, ε, options)
    return σ, tan, new


def _response_plastic(m, ε, state, options, want_tangent, out=None):
    layout = state.layout
    ε = _prep(ε)
    N = _npoints(ε, 6, layout)
    if N != state.N:
        raise ValueError("strain has %d points, state has %d" % (N, state.N))
    sin = _prep(state.data)
    _same_place(ε, sin, "material_response(::Plastic)")
    host = _is_host(ε)
    σ = _out(out, "stress", _shape(6, N, layout), ε)
    tan = _out(out, "tangent", _shape(36, N, layout), ε) if want_tangent else None
    sout = _out(out, "state", _shape(14, N, layout), ε)
    flags = _out(out, "flags", (N,), ε, np.uint8 if host else _torch().uint8)
    iters = _out(out, "iters", (N,), ε, np.uint16 if host else _torch().int16)
    nfail = _zeros((1,), ε, np.int32 if host else _torch().int32)
    o = _newton_opts(m, options)
    lib = load()
    with _DeviceGuard(ε):
        if host:
            rc = lib.mmh_plastic(ctypes.byref(m._c), _ptr(ε), _ptr(sin), _ptr(σ), _ptr(tan), _ptr(sout), _ptr(flags), _ptr(iters),
                                 _ptr(nfail), ctypes.byref(o), N, _LAYOUTS[layout])
        else:
            rc = lib.mm_plastic(ctypes.byref(m._c), _ptr(ε), _ptr(sin), _ptr(σ), _ptr(tan), _ptr(sout), _ptr(flags), _ptr(iters),
                                _ptr(nfail), ctypes.byref(o), N, _LAYOUTS[layout], _stream(ε))
    check(rc, "mm_plastic")
    new = PlasticState(sout, layout)
    new.flags, new.iters, new.n_fail = flags, iters, nfail
    _finish_fail(nfail, ε, options)
    return σ, tan, new


def _response_crystal(m, Δε, state, Δt, options, want_tangent, out=None):
    layout = state.layout
    Δε = _prep(Δε)
    N = _npoints(Δε, 6, layout)
    if N != state.N:
        raise ValueError("strain has %d points, state has %d" % (N, state.N))
    sin = _prep(state.data)
    if _is_host(Δε) and not _is_host(sin) and options and options.get("state_resident"):
        # opt-in (options={"state_resident": True}): unlike the reference's call this updates the state array IN PLACE
        return _response_crystal_resident(m, Δε, state, sin, Δt, options, want_tangent, out, N, layout)
    _same_place(Δε, sin, "material_response(::CrystalViscoPlastic)")
    σ = _out(out, "stress", _shape(6, N, layout), Δε)
    tan = _out(out, "tangent", _shape(36, N, layout), Δε) if want_tangent else None
    sout = _out(out, "state", _shape(42, N, layout), Δε)
    host = _is_host(Δε)
    active = _out(out, "flags", (N,), Δε, np.uint16 if host else _torch().int16)
    iters = _out(out, "iters", (N,), Δε, np.uint16 if host else _torch().int16)
    nfail = _zeros((1,), Δε, np.int32 if host else _torch().int32)
    o = _newton_opts(m, options)
    lib = load()
    with _DeviceGuard(Δε):
        if host:
            rc = lib.mmh_crystal(ctypes.byref(m._c), _ptr(Δε), float(Δt), _ptr(sin), _ptr(σ), _ptr(tan), _ptr(sout), _ptr(active),
                                 _ptr(iters), _ptr(nfail), ctypes.byref(o), N, _LAYOUTS[layout])
        else:
            rc = lib.mm_crystal(ctypes.byref(m._c), _ptr(Δε), float(Δt), _ptr(sin), _ptr(σ), _ptr(tan), _ptr(sout), _ptr(active),
                                _ptr(iters), _ptr(nfail), ctypes.byref(o), N, _LAYOUTS[layout], _stream(Δε))
    check(rc, "mm_crystal")
    new = CrystalViscoPlasticState(sout, layout)
    new.flags, new.iters, new.n_fail = active, iters, nfail
    _finish_fail(nfail, Δε, options)
    return σ, tan, new


def _response_crystal_resident(m, Δε, state, sdev, Δt, options, want_tangent, out, N, layout):
    """options={"state_resident": True}: host strains, state resident on a CUDA device (what a time-stepping caller keeps between steps):
    the state array is updated IN PLACE and handed back; only Δε crosses the link on the way up, σ / ∂σ∂ε / masks / counts on the way down (mmh_crystal_resident)."""
    if out is not None and out.get("state") is not None:
        raise ValueError("material_response(::CrystalViscoPlastic): the device-resident state is updated in place; out['state'] is not used")
    σ = _out(out, "stress", _shape(6, N, layout), Δε)
    tan = _out(out, "tangent", _shape(36, N, layout), Δε) if want_tangent else None
    active = _out(out, "flags", (N,), Δε, np.uint16)
    iters = _out(out, "iters", (N,), Δε, np.uint16)
    nfail = _zeros((1,), Δε, np.int32)
    o = _newton_opts(m, options)
    with _DeviceGuard(sdev):
        rc = load().mmh_crystal_resident(ctypes.byref(m._c), _ptr(Δε), float(Δt), _ptr(sdev), _ptr(σ), _ptr(tan), _ptr(active), _ptr(iters),
                                         _ptr(nfail), ctypes.byref(o), N, _LAYOUTS[layout])
    check(rc, "mmh_crystal_resident")
    new = CrystalViscoPlasticState(sdev, layout)
    new.flags, new.iters, new.n_fail = active, iters, nfail
    _finish_fail(nfail, Δε, options)
    return σ, tan, new


def _response_hyper(m, strain, kind, state, layout, want_tangent, tangent_kind=_abi.MM_TANGENT_DSDC, out=None):
    strain = _prep(strain)
    nin = 9 if kind == _abi.MM_STRAIN_F else 6
    ns, nt = (9, 81) if tangent_kind == _abi.MM_TANGENT_DPDF else (6, 36)
    N = _npoints(strain, nin, layout)
    S = _out(out, "stress", _shape(ns, N, layout), strain)
    tan = _out(out, "tangent", _shape(nt, N, layout), strain) if want_tangent else None
    lib = load()
    with _DeviceGuard(strain):
        if _is_host(strain):
            rc = lib.mmh_hyper_ex(m.MODEL, ctypes.byref(m._c), _ptr(strain), kind, tangent_kind, _ptr(S), _ptr(tan), N, _LAYOUTS[layout])
        else:
            rc = lib.mm_hyper_ex(m.MODEL, ctypes.byref(m._c), _ptr(strain), kind, tangent_kind, _ptr(S), _ptr(tan), N, _LAYOUTS[layout], _stream(strain))
    check(rc, "mm_hyper")
    return S, tan, state


def _response_transviso(m, ε, state, want_tangent, out=None):
    """material_response(m::TransverselyIsotropic, ε, state) — reference src/transverselyisotropic.jl:92-106."""
    layout = state.layout
    ε = _prep(ε)
    N = _npoints(ε, 6, layout)
    if N != state.N:
        raise ValueError("strain has %d points, state has %d" % (N, state.N))
    a3 = _prep(state.data)
    if _is_host(ε) != _is_host(a3):
        raise ValueError("strain and state must live in the same place")
    σ = _out(out, "stress", _shape(6, N, layout), ε)
    tan = _out(out, "tangent", _shape(36, N, layout), ε) if want_tangent else None
    lib = load()
    with _DeviceGuard(ε):
        if _is_host(ε):
            rc = lib.mmh_transversely_isotropic(ctypes.byref(m._c), _ptr(ε), _ptr(a3), _ptr(σ), _ptr(tan), N, _LAYOUTS[layout])
        else:
            rc = lib.mm_transversely_isotropic(ctypes.byref(m._c), _ptr(ε), _ptr(a3), _ptr(σ), _ptr(tan), N, _LAYOUTS[layout], _stream(ε))
    check(rc, "mm_transversely_isotropic")
    return σ, tan, state


def elastic_strain_energy_density(m, strain, layout="soa"):
    """elastic_strain_energy_density(m, strain) -> ψ[N] — reference src/LinearElastic.jl:63 (ε), neohook.jl:29-33,
    yeoh.jl:31-35, stvenant.jl:28-34 (C), and with a StrainMeasure (C / F / E) through src/traits.jl:133-136."""
    kind = _abi.MM_STRAIN_C
    if isinstance(strain, StrainMeasure):
        if isinstance(m, _Hyper):
            if isinstance(strain, DeformationGradient):
                kind = _abi.MM_STRAIN_F
            elif isinstance(strain, GreenLagrange):
                kind = _abi.MM_STRAIN_E
            elif not isinstance(strain, RightCauchyGreen):
                raise RuntimeError("%s is not the native strain measure for %s:" % (type(strain).__name__, type(m).__name__))
        elif not isinstance(strain, SmallStrain):
            raise RuntimeError("%s is not the native strain measure for %s:" % (type(strain).__name__, type(m).__name__))
        strain = strain.value
    x = _prep(strain)
    lib = load()
    if isinstance(m, LinearElastic):
        N = _npoints(x, 6, layout)
        ψ = _empty((N,), x)
        with _DeviceGuard(x):
            if _is_host(x):
                rc = lib.mmh_linear_elastic_energy(ctypes.byref(m._c), _ptr(x), _ptr(ψ), N, _LAYOUTS[layout])
            else:
                rc = lib.mm_linear_elastic_energy(ctypes.byref(m._c), _ptr(x), _ptr(ψ), N, _LAYOUTS[layout], _stream(x))
        check(rc, "mm_linear_elastic_energy")
        return ψ
    if isinstance(m, _Hyper):
        N = _npoints(x, 9 if kind == _abi.MM_STRAIN_F else 6, layout)
        ψ = _empty((N,), x)
        with _DeviceGuard(x):
            if _is_host(x):
                rc = lib.mmh_hyper_energy(m.MODEL, ctypes.byref(m._c), _ptr(x), kind, _ptr(ψ), N, _LAYOUTS[layout])
            else:
                rc = lib.mm_hyper_energy(m.MODEL, ctypes.byref(m._c), _ptr(x), kind, _ptr(ψ), N, _LAYOUTS[layout], _stream(x))
        check(rc, "mm_hyper_energy")
        return ψ
    raise TypeError("elastic_strain_energy_density is defined for LinearElastic, NeoHook, Yeoh, StVenant (as in the reference)")


def _response_xu(m, Δ, state, layout, want_tangent, out=None):
    Δ = _prep(Δ)
    dim = Δ.shape[0] if _LAYOUTS[layout] == 0 else Δ.shape[1]
    if dim not in (2, 3):
        raise ValueError("XuNeedleman: jump must have 2 or 3 components")
    N = _npoints(Δ, dim, layout)
    T = _out(out, "stress", _shape(dim, N, layout), Δ)
    H = _out(out, "tangent", _shape(dim * dim, N, layout), Δ) if want_tangent else None
    lib = load()
    with _DeviceGuard(Δ):
        if _is_host(Δ):
            rc = lib.mmh_xu_needleman(ctypes.byref(m._c), _ptr(Δ), _ptr(T), _ptr(H), dim, N, _LAYOUTS[layout])
        else:
            rc = lib.mm_xu_needleman(ctypes.byref(m._c), _ptr(Δ), _ptr(T), _ptr(H), dim, N, _LAYOUTS[layout], _stream(Δ))
    check(rc, "mm_xu_needleman")
    return T, H, state


def _response_wrapped(dim, m, Δε, state, options, layout, want_tangent, Δt=None):
    """material_response(dim::AbstractDim, m, Δε, state, Δt; cache, options) — reference src/wrappers.jl:29-79
    (LinearElastic, Plastic, TransverselyIsotropic, CrystalViscoPlastic; device arrays -> mm_wrapped_*, numpy arrays -> mmh_wrapped_*)."""
    kind = dim.KIND
    lib = load()
    nc = lib.mm_wrap_ncomp(kind)
    wtol, wmax = 1e-8, 10                                            # wrappers.jl:55-56
    if options:
        wtol = float(options.get("plane_stress_tol", options.get(":plane_stress_tol", wtol)))
        wmax = int(options.get("plane_stress_max_iter", options.get(":plane_stress_max_iter", wmax)))
    w = _abi.MMNewtonOpts(wtol, wmax, 0)
    if isinstance(m, (Plastic, TransverselyIsotropic, CrystalViscoPlastic)):
        if state is None:
            raise ValueError("material_response(dim, ::%s, Δε, state): state required" % type(m).__name__)
        layout = state.layout
    if isinstance(m, CrystalViscoPlastic) and Δt is None:
        raise ValueError("material_response(dim, ::CrystalViscoPlastic, Δε, state, Δt): Δt required")
    Δε = _prep(Δε)
    host = _is_host(Δε)
    N = _npoints(Δε, nc, layout)
    u8 = np.uint8 if host else _torch().uint8
    σ = _empty(_shape(nc, N, layout), Δε)
    tan = _empty(_shape(nc * nc, N, layout), Δε) if want_tangent else None
    oit = _empty((N,), Δε, u8)
    nfail = _zeros((1,), Δε, np.int32 if host else _torch().int32)
    defer = bool(options and options.get("defer_check"))
    lay = _LAYOUTS[layout]
    tail = () if host else (_stream(Δε),)
    flags = None
    with _DeviceGuard(Δε):
        if isinstance(m, LinearElastic):
            fn = lib.mmh_wrapped_linear_elastic if host else lib.mm_wrapped_linear_elastic
            check(fn(kind, ctypes.byref(m._c), _ptr(Δε), _ptr(σ), _ptr(tan), _ptr(oit), _ptr(nfail), ctypes.byref(w), N, lay, *tail), "mm_wrapped_linear_elastic")
            new = state if state is not None else LinearElasticState()
        elif isinstance(m, TransverselyIsotropic):
            if N != state.N:
                raise ValueError("strain has %d points, state has %d" % (N, state.N))
            a3 = _prep(state.data)
            if _is_host(a3) != host:
                raise ValueError("strain and state must live in the same place")
            fn = lib.mmh_wrapped_transversely_isotropic if host else lib.mm_wrapped_transversely_isotropic
            check(fn(kind, ctypes.byref(m._c), _ptr(Δε), _ptr(a3), _ptr(σ), _ptr(tan), _ptr(oit), _ptr(nfail), ctypes.byref(w), N, lay, *tail),
                  "mm_wrapped_transversely_isotropic")
            new = state
        elif isinstance(m, Plastic):
            if N != state.N:
                raise ValueError("material_response(dim, ::Plastic, Δε, state): state with %d points required" % N)
            sin = _prep(state.data)
            if _is_host(sin) != host:
                raise ValueError("strain and state must live in the same place")
            sout = _empty(_shape(14, N, layout), Δε)
            flags = _empty((N,), Δε, u8)
            iters = _empty((N,), Δε, np.uint16 if host else _torch().int16)
            o = _newton_opts(m, options)
            fn = lib.mmh_wrapped_plastic if host else lib.mm_wrapped_plastic
            check(fn(kind, ctypes.byref(m._c), _ptr(Δε), _ptr(sin), _ptr(σ), _ptr(tan), _ptr(sout), _ptr(flags), _ptr(iters), _ptr(oit), _ptr(nfail),
                     ctypes.byref(o), ctypes.byref(w), N, lay, *tail), "mm_wrapped_plastic")
            new = PlasticState(sout, layout)
            new.flags, new.iters, new.n_fail = flags, iters, nfail
        elif isinstance(m, CrystalViscoPlastic):
            if N != state.N:
                raise ValueError("strain has %d points, state has %d" % (N, state.N))
            sin = _prep(state.data)
            if _is_host(sin) != host:
                raise ValueError("strain and state must live in the same place")
            sout = _empty(_shape(42, N, layout), Δε)
            active = _empty((N,), Δε, np.uint16 if host else _torch().int16)
            iters = _empty((N,), Δε, np.uint16 if host else _torch().int16)
            o = _newton_opts(m, options)
            fn = lib.mmh_wrapped_crystal if host else lib.mm_wrapped_crystal
            check(fn(kind, ctypes.byref(m._c), _ptr(Δε), float(Δt), _ptr(sin), _ptr(σ), _ptr(tan), _ptr(sout), _ptr(active), _ptr(iters), _ptr(oit),
                     _ptr(nfail), ctypes.byref(o), ctypes.byref(w), N, lay, *tail), "mm_wrapped_crystal")
            new = CrystalViscoPlasticState(sout, layout)
            new.flags, new.iters, new.n_fail = active, iters, nfail
            flags = ((active >> 8) & 0x20) << 1                       # MM_ACTIVE_WRAPPER_FAILED (bit 13) -> MM_FLAG_WRAPPER_FAILED (0x40)
        else:
            raise NotImplementedError("dimension wrappers are implemented for LinearElastic, Plastic, TransverselyIsotropic and CrystalViscoPlastic")
    if hasattr(new, "__dict__"):
        new.outer_iters = oit
        new.n_fail = nfail
    if not defer and (int(nfail[0]) if host else int(nfail.item())) != 0:
        if flags is None or bool((flags & _abi.MM_FLAG_WRAPPER_FAILED).any()):
            raise RuntimeError(PLANE_STRESS_NOT_CONVERGED)
        raise RuntimeError(NOT_CONVERGED)
    return σ, tan, new


def material_response(*args, cache=None, options=None, layout="soa", tangent=True, out=None):
    """material_response(m, strain, state=None, Δt=None, cache=None, options=None)
    material_response(dSdC(), m, DeformationGradient(F) | RightCauchyGreen(C), state=None, Δt=0.0)

    Returns (stress, tangent, new_state) for all N points of `strain`.
    `layout` applies to stateless models (stateful ones use their state's layout);
    `tangent=False` skips the 36-component tangent output (returns None in its place);
    `tangent="compact" | "factored"` (Plastic, numpy strains): the reference's own return shape, src/Plastic.jl:134-135 — elastic points
    share m.Eᵉ and keep their state, yielded points own one row (CompactTangent / CompactPlasticState); a state on a CUDA device
    with numpy strains = resident mode (state updated in place on the device, only ε up and σ + rows down);
    `out` = dict of preallocated output buffers (e.g. pinned host memory): stress, tangent [, state, flags, iters for Plastic / CrystalViscoPlastic; rows for the compact forms]."""
    args = list(args)
    if args and isinstance(args[0], AbstractDim):                            # src/wrappers.jl:29-79
        dim = args.pop(0)
        if dim.KIND == 0:                                                    # Dim{3}: the plain 3D call (wrappers.jl:22, :29-41)
            return material_response(*args, cache=cache, options=options, layout=layout, tangent=tangent, out=out)
        m = args.pop(0)
        strain = args.pop(0)
        state = args.pop(0) if args else None
        Δt = args.pop(0) if args else None
        return _response_wrapped(dim, m, strain, state, options, layout, tangent, Δt)
    if args and isinstance(args[0], AbstractTangent):                       # src/traits.jl:91-107
        out_tangent = args.pop(0)
        m = args.pop(0)
        strain = args.pop(0)
        state = args.pop(0) if args else None
        if not isinstance(strain, StrainMeasure):
            raise TypeError("trait form expects a StrainMeasure (DeformationGradient / RightCauchyGreen / GreenLagrange / SmallStrain)")
        if not isinstance(m, _Hyper):
            # small-strain materials: native strain SmallStrain, native tangent ∂σ∂ε (traits.jl:142-143, LinearElastic.jl:25); both
            # transformations are the identity (traits.jl:33, :58), so this is the native 3D call with the wrapper taken off
            if not isinstance(m, (LinearElastic, Plastic, TransverselyIsotropic, CrystalViscoPlastic)):
                raise TypeError("material_response(tangent, m, strain, ...): unsupported material %r" % (m,))
            if not isinstance(strain, SmallStrain):                          # traits.jl:120
                raise RuntimeError("%s is not the native strain measure for %s:" % (type(strain).__name__, type(m).__name__))
            if not isinstance(out_tangent, dsigmadeps):                      # no transform_stress_tangent method from ∂σ∂ε to anything else (traits.jl:58-77)
                raise TypeError("no transformation from dsigmadeps (∂σ∂ε) to %s is defined (src/traits.jl:58-77)" % type(out_tangent).__name__)
            rest = [state] + args if state is not None else []
            return material_response(m, strain.value, *rest, cache=cache, options=options, layout=layout, tangent=tangent, out=out)
        if isinstance(out_tangent, dsigmadeps):
            raise TypeError("no transformation from dSdC (∂S∂C) to dsigmadeps is defined (src/traits.jl:58-77)")
        if isinstance(out_tangent, dSdC):
            tk = _abi.MM_TANGENT_DSDC
        elif isinstance(out_tangent, dSdE):
            tk = _abi.MM_TANGENT_DSDE
        elif isinstance(out_tangent, dPdF):
            tk = _abi.MM_TANGENT_DPDF
        else:
            raise NotImplementedError("%s output is not defined for %s" % (type(out_tangent).__name__, type(m).__name__))
        if isinstance(strain, DeformationGradient):
            kind = _abi.MM_STRAIN_F
        elif isinstance(strain, RightCauchyGreen):
            kind = _abi.MM_STRAIN_C
        elif isinstance(strain, GreenLagrange):
            kind = _abi.MM_STRAIN_E                                          # traits.jl:43-46
        else:                                                                # traits.jl:120
            raise RuntimeError("%s is not the native strain measure for %s:" % (type(strain).__name__, type(m).__name__))
        if tk == _abi.MM_TANGENT_DPDF and kind != _abi.MM_STRAIN_F:
            raise TypeError("the dPdF (∂Pᵀ∂F) output needs DeformationGradient(F) as input (traits.jl:67-72)")
        return _response_hyper(m, strain.value, kind, state if state is not None else initial_material_state(m), layout, tangent, tk, out)

    m = args.pop(0)
    strain = args.pop(0)
    state = args.pop(0) if args else None
    Δt = args.pop(0) if args else None
    if args:                                   # XuNeedleman takes cache, options positionally (XuNeedleman.jl:75-76)
        cache = args.pop(0)
    if args:
        options = args.pop(0)
    if isinstance(m, LinearElastic):
        return _response_linear_elastic(m, strain, state if state is not None else LinearElasticState(), layout, tangent, out)
    if isinstance(m, Plastic):
        if state is None:
            raise TypeError("material_response(::Plastic, ε, state): state required")
        if isinstance(tangent, str):
            if tangent not in ("compact", "factored"):
                raise ValueError("tangent must be True, False, 'compact' or 'factored'")
            return _response_plastic_compact(m, strain, state, options, tangent, out)
        return _response_plastic(m, strain, state, options, tangent, out)
    if isinstance(m, CrystalViscoPlastic):
        if state is None:
            raise TypeError("material_response(::CrystalViscoPlastic, Δε, state): state required")
        return _response_crystal(m, strain, state, 1.0 if Δt is None else Δt, options, tangent, out)
    if isinstance(m, _Hyper):
        return _response_hyper(m, strain, _abi.MM_STRAIN_C, state if state is not None else initial_material_state(m), layout, tangent, out=out)
    if isinstance(m, TransverselyIsotropic):
        if state is None:
            raise TypeError("material_response(::TransverselyIsotropic, ε, state): state (material direction) required")
        return _response_transviso(m, strain, state, tangent, out)
    if isinstance(m, XuNeedleman):
        return _response_xu(m, strain, state if state is not None else XuNeedlemanState(), layout, tangent, out)
    raise TypeError("material_response: unsupported material %r" % (m,))


def cell_internal_forces(m, dNdx, dΩ, ue, state=None, tangent=False, stress=False, stiffness=False, options=None):
    """The caller's quadrature loop fused around material_response (what a Ferrite.jl-style
    assemble_cell! does around the reference's per-point call, src/MaterialModels.jl:41-52), all cells at once:

        ε = sym(Σ_a u_a ⊗ ∇N_a);   σ, D, state_new = material_response(m, ε, state);   fe[a] += (σ ⋅ ∇N_a) dΩ

    dNdx (3·nn, P): ∇N_a at every quadrature point (physical coordinates), dΩ (P,), ue (3·nn, ncell), P = ncell·nq, point
    index = cell·nq + q; state: PlasticState over the P points (layout "soa") for Plastic, None for LinearElastic.
    Returns (fe (3·nn, ncell), new_state, extras) with extras["stress"] (6, P) / extras["tangent"] (36, P) when requested and
    extras["stiffness"] ((3·nn)², ncell): ke[3a+i, 3b+j] = Σ_q δε ⊡ D ⊡ Δε dΩ at row (3a+i)·3nn + 3b+j (stiffness=True).
    NeoHook / Yeoh / StVenant: total-Lagrangian form, F = I + Σ_a u_a ⊗ ∇N_a, fe[a] = Σ (P⋅∇N_a) dΩ, extras["stress"] = Pᵀ (9, P),
    extras["tangent"] = ∂Pᵀ∂F (81, P).  Device (torch CUDA) arrays only.  Supported (nn, nq): (4,1) (4,4) (8,1) (8,8) (10,4)."""
    dNdx, dΩ, ue = _prep(dNdx), _prep(dΩ), _prep(ue)
    if _is_host(dNdx) or _is_host(dΩ) or _is_host(ue):
        raise NotImplementedError("cell_internal_forces: device (torch CUDA) arrays only")
    if ue.dim() != 2 or ue.shape[0] % 3 != 0:
        raise ValueError("ue must have shape (3*nn, ncell)")
    nn, ncell = ue.shape[0] // 3, ue.shape[1]
    P = dΩ.shape[0]
    if ncell == 0 or P % ncell != 0 or tuple(dNdx.shape) != (3 * nn, P):
        raise ValueError("inconsistent shapes: dNdx %s, dΩ %s, ue %s" % (tuple(dNdx.shape), tuple(dΩ.shape), tuple(ue.shape)))
    nq = P // ncell
    torch = _torch()
    fe = _empty((3 * nn, ncell), ue)
    σ = _empty((6, P), ue) if stress else None
    tan = _empty((36, P), ue) if tangent else None
    ke = _empty((9 * nn * nn, ncell), ue) if stiffness else None
    lib = load()
    extras = {"stress": σ, "tangent": tan, "stiffness": ke}
    with _DeviceGuard(ue):
        if isinstance(m, LinearElastic):
            check(lib.mm_cells_linear_elastic(ctypes.byref(m._c), nn, nq, _ptr(dNdx), _ptr(dΩ), _ptr(ue), _ptr(fe), _ptr(ke), _ptr(σ), _ptr(tan), ncell,
                                              _stream(ue)), "mm_cells_linear_elastic")
            return fe, (state if state is not None else LinearElasticState()), extras
        if isinstance(m, TransverselyIsotropic):
            if state is None or state.N != P or _LAYOUTS[state.layout] != _abi.MM_LAYOUT_SOA:
                raise ValueError("cell_internal_forces(::TransverselyIsotropic): a 'soa' TransverselyIsotropicState over the %d quadrature points is required" % P)
            check(lib.mm_cells_transversely_isotropic(ctypes.byref(m._c), nn, nq, _ptr(dNdx), _ptr(dΩ), _ptr(ue), _ptr(_prep(state.data)), _ptr(fe), _ptr(ke),
                                                      _ptr(σ), _ptr(tan), ncell, _stream(ue)), "mm_cells_transversely_isotropic")
            return fe, state, extras
        if isinstance(m, _Hyper):                       # total Lagrangian: F = I + Σ u ⊗ ∇N, stress = Pᵀ (9), tangent = ∂Pᵀ∂F (81)
            if stiffness:
                raise NotImplementedError("cell_internal_forces(::%s): element stiffness not implemented (ask for tangent=True and assemble from ∂Pᵀ∂F)" % type(m).__name__)
            Pt = _empty((9, P), ue) if stress else None
            dP = _empty((81, P), ue) if tangent else None
            check(lib.mm_cells_hyper(m.MODEL, ctypes.byref(m._c), nn, nq, _ptr(dNdx), _ptr(dΩ), _ptr(ue), _ptr(fe), _ptr(Pt), _ptr(dP), ncell, _stream(ue)),
                  "mm_cells_hyper")
            return fe, (state if state is not None else initial_material_state(m)), {"stress": Pt, "tangent": dP, "stiffness": None}
        if not isinstance(m, Plastic):
            raise NotImplementedError("cell_internal_forces is implemented for LinearElastic, Plastic, NeoHook, Yeoh, StVenant")
        if state is None or state.N != P or _LAYOUTS[state.layout] != _abi.MM_LAYOUT_SOA:
            raise ValueError("cell_internal_forces(::Plastic): a 'soa' PlasticState over the %d quadrature points is required" % P)
        sin = _prep(state.data)
        _same_place(ue, sin, "cell_internal_forces(::Plastic)")
        sout = _empty((14, P), ue)
        flags = _empty((P,), ue, torch.uint8)
        iters = _empty((P,), ue, torch.int16)
        nfail = _zeros((1,), ue, torch.int32)
        o = _newton_opts(m, options)
        check(lib.mm_cells_plastic(ctypes.byref(m._c), nn, nq, _ptr(dNdx), _ptr(dΩ), _ptr(ue), _ptr(sin), _ptr(fe), _ptr(ke), _ptr(σ), _ptr(tan), _ptr(sout), _ptr(flags),
                                   _ptr(iters), _ptr(nfail), ctypes.byref(o), ncell, _stream(ue)), "mm_cells_plastic")
    new = PlasticState(sout, "soa")
    new.flags, new.iters, new.n_fail = flags, iters, nfail
    _finish_fail(nfail, ue, options)
    return fe, new, extras


def synth_uniform(N, nc, seed, lo, hi, layout="soa", device="cuda"):
    """Device-side deterministic inputs (bit-identical to synth.host_uniform)."""
    torch = _torch()
    out = torch.empty(_shape(nc, N, layout), dtype=torch.float64, device=device)
    lo = np.ascontiguousarray(np.broadcast_to(np.asarray(lo, dtype=np.float64), (nc,)))
    hi = np.ascontiguousarray(np.broadcast_to(np.asarray(hi, dtype=np.float64), (nc,)))
    with _DeviceGuard(out):
        rc = load().mm_synth_uniform(_ptr(out), N, nc, _LAYOUTS[layout], int(seed), _ptr(lo), _ptr(hi), _stream(out))
    check(rc, "mm_synth_uniform")
    return out


class _PinnedBlock:
    """owner of one mmh_alloc_pinned allocation; freed when the last numpy view that refers to it is gone"""

    def __init__(self, nbytes):
        self.ptr = load().mmh_alloc_pinned(max(nbytes, 1))
        if not self.ptr:
            raise MemoryError("mmh_alloc_pinned(%d) failed" % nbytes)
        self.nbytes = max(nbytes, 1)
        self.__array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def __del__(self):
        try:
            if self.ptr:
                load().mmh_free_pinned(ctypes.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


class PinnedArray:
    """float64/uint8/... numpy array backed by page-locked host memory (mmh_alloc_pinned): what a host-resident caller should hand
    to the numpy path so the H2D/D2H copies run at PCIe speed and truly asynchronously.  `.array` and every view or slice of it keep
    the allocation alive (the pinned block is the numpy base object), so `PinnedArray(shape).array` is safe on its own and results
    returned by material_response(out=...) stay valid after the PinnedArray object is gone; the memory is unpinned and released when
    the last view dies.  `free()` only drops this object's own reference."""

    def __init__(self, shape, dtype=np.float64):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        block = _PinnedBlock(self.nbytes)
        raw = np.asarray(block)                      # uint8 view whose .base is the block: numpy's refcount owns the pinned memory
        self.array = raw[: self.nbytes].view(dtype).reshape(shape)

    def free(self):
        self.array = None
