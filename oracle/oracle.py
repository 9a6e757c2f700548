"""ORACLE — TEST INFRASTRUCTURE ONLY.

numpy/ctypes binding of oracle/libmm_oracle.so (the CPU restatement of the reference's
`material_response` methods, see mmo_models.hpp).  May be imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; the product package
never imports it.

Arrays are float64 numpy arrays: SoA = shape (nc, N) C-contiguous, AoS = shape (N, nc) C-contiguous.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import materialmodels_jl_b200._abi as abi  # noqa: E402  (POD structs only; no CUDA involved)

SOA, AOS = abi.MM_LAYOUT_SOA, abi.MM_LAYOUT_AOS
_P, _I64 = C.c_void_p, C.c_int64

_SIG = {
    "mmo_num_threads": (C.c_int, []),
    "mmo_set_threads": (None, [C.c_int]),
    "mmo_elastic_tangent_3d": (C.c_int, [C.c_double, C.c_double, _P]),
    "mmo_slipsystems": (C.c_int, [C.c_int, _P, _P, _P]),
    "mmo_crystal_setup": (C.c_int, [C.POINTER(abi.MMCrystal), _P, _P]),
    "mmo_xu_needleman_setup": (C.c_int, [C.c_double] * 5 + [C.POINTER(abi.MMXuNeedleman)]),
    "mmo_linear_elastic": (C.c_int, [C.POINTER(abi.MMLinearElastic), _P, _P, _P, _I64, C.c_int]),
    "mmo_plastic": (C.c_int, [C.POINTER(abi.MMPlastic), _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(abi.MMNewtonOpts), _I64, C.c_int]),
    "mmo_hyper": (C.c_int, [C.c_int, C.POINTER(abi.MMHyper), _P, C.c_int, _P, _P, _I64, C.c_int]),
    "mmo_hyper_ad": (C.c_int, [C.c_int, C.POINTER(abi.MMHyper), _P, _P, _P, _P]),
    "mmo_crystal": (C.c_int, [C.POINTER(abi.MMCrystal), _P, C.c_double, _P, _P, _P, _P, _P, _P, _P, C.POINTER(abi.MMNewtonOpts), _I64, C.c_int]),
    "mmo_xu_needleman": (C.c_int, [C.POINTER(abi.MMXuNeedleman), _P, _P, _P, C.c_int, _I64, C.c_int]),
    "mmo_solve_scalar_kat": (C.c_int, [C.c_int, C.c_double, _P, _P, _P]),
    "mmo_strain_energy": (C.c_double, [C.c_int, C.POINTER(abi.MMHyper), _P]),
    "mmo_synth_uniform": (C.c_int, [_P, _I64, C.c_int, C.c_int, C.c_uint64, _P, _P]),
    "mmo_hyper_ex": (C.c_int, [C.c_int, C.POINTER(abi.MMHyper), _P, C.c_int, C.c_int, _P, _P, _I64, C.c_int]),
    "mmo_neohook_dPdF_ad": (C.c_int, [C.POINTER(abi.MMHyper), _P, _P, _P]),
    "mmo_wrapped_transviso": (C.c_int, [C.c_int, C.POINTER(abi.MMTransverselyIsotropic), _P, _P, _P, _P, _P, _P, C.POINTER(abi.MMNewtonOpts), _I64, C.c_int]),
    "mmo_cells": (C.c_int, [C.c_int, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(abi.MMNewtonOpts), _I64]),
    "mmo_cells_hyper": (C.c_int, [C.c_int, C.POINTER(abi.MMHyper), C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _I64]),
    "mmo_wrapped_crystal": (C.c_int, [C.c_int, C.POINTER(abi.MMCrystal), _P, C.c_double, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(abi.MMNewtonOpts),
                                      C.POINTER(abi.MMNewtonOpts), _I64, C.c_int]),
    "mmo_transviso_setup": (C.c_int, [C.c_double] * 5 + [C.POINTER(abi.MMTransverselyIsotropic)]),
    "mmo_transversely_isotropic": (C.c_int, [C.POINTER(abi.MMTransverselyIsotropic), _P, _P, _P, _P, _I64, C.c_int]),
    "mmo_strain_energy_batch": (C.c_int, [C.c_int, _P, _P, C.c_int, _P, _I64, C.c_int]),
    "mmo_wrapped_linear_elastic": (C.c_int, [C.c_int, C.POINTER(abi.MMLinearElastic), _P, _P, _P, _P, _P, C.POINTER(abi.MMNewtonOpts), _I64, C.c_int]),
    "mmo_wrapped_plastic": (C.c_int, [C.c_int, C.POINTER(abi.MMPlastic), _P, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(abi.MMNewtonOpts),
                                      C.POINTER(abi.MMNewtonOpts), _I64, C.c_int]),
}

_lib = None


def build(force=False):
    so = os.path.join(_HERE, "libmm_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("mm_oracle.cpp", "mmo_models.hpp", "mmo_tensor.hpp")]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "mmb200.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libmm_oracle.so"])
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        for name, (res, args) in _SIG.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _alloc(nc, N, layout, dtype=np.float64):
    return np.zeros((nc, N) if layout == SOA else (N, nc), dtype=dtype)


def _n(a, layout):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, (a.shape[1] if layout == SOA else a.shape[0])


def num_threads():
    return lib().mmo_num_threads()


def set_threads(n):
    lib().mmo_set_threads(int(n))


def elastic_tangent_3d(E, nu):
    out = np.zeros(36)
    lib().mmo_elastic_tangent_3d(E, nu, _ptr(out))
    return out


def slipsystems(structure, rodrigues):
    r = np.asarray(rodrigues, dtype=np.float64)
    planes, dirs = np.zeros((12, 3)), np.zeros((12, 3))
    n = lib().mmo_slipsystems(structure, _ptr(r), _ptr(planes), _ptr(dirs))
    assert n == 12
    return planes, dirs


def crystal_params(E, nu, tau_y, H_iso, H_kin, q, alpha_inf, t_star, sigma_c, m, planes, dirs):
    p = abi.MMCrystal(E=E, nu=nu, tau_y=tau_y, H_iso=H_iso, H_kin=H_kin, q=q, alpha_inf=alpha_inf, t_star=t_star, sigma_c=sigma_c, m=m)
    planes = np.ascontiguousarray(planes, dtype=np.float64)
    dirs = np.ascontiguousarray(dirs, dtype=np.float64)
    lib().mmo_crystal_setup(C.byref(p), _ptr(planes), _ptr(dirs))
    return p


def xu_params(sigma_max, tau_max, Phi_n, Phi_t, Delta_n_star):
    p = abi.MMXuNeedleman()
    lib().mmo_xu_needleman_setup(sigma_max, tau_max, Phi_n, Phi_t, Delta_n_star, C.byref(p))
    return p


def linear_elastic(E, nu, eps, layout=SOA, want_tangent=True):
    eps, N = _n(eps, layout)
    sig = _alloc(6, N, layout)
    tan = _alloc(36, N, layout) if want_tangent else None
    p = abi.MMLinearElastic(E, nu)
    rc = lib().mmo_linear_elastic(C.byref(p), _ptr(eps), _ptr(sig), _ptr(tan), N, layout)
    assert rc == 0
    return sig, tan


def plastic(params, eps, state, layout=SOA, tol=1e-8, max_iter=1000):
    """params: abi.MMPlastic or tuple (E, nu, sigma_y, H, r, kappa_inf, alpha_inf). Returns dict."""
    p = params if isinstance(params, abi.MMPlastic) else abi.MMPlastic(*params)
    eps, N = _n(eps, layout)
    state = np.ascontiguousarray(state, dtype=np.float64)
    sig, tan, st = _alloc(6, N, layout), _alloc(36, N, layout), _alloc(14, N, layout)
    flags, iters = np.zeros(N, np.uint8), np.zeros(N, np.uint16)
    nfail = np.zeros(1, np.int32)
    o = abi.MMNewtonOpts(tol, max_iter, 0)
    rc = lib().mmo_plastic(C.byref(p), _ptr(eps), _ptr(state), _ptr(sig), _ptr(tan), _ptr(st), _ptr(flags), _ptr(iters), _ptr(nfail), C.byref(o), N, layout)
    assert rc == 0
    return dict(sig=sig, tan=tan, state=st, flags=flags, iters=iters, n_fail=int(nfail[0]))


def hyper(model, params, strain, strain_kind, layout=SOA):
    p = params if isinstance(params, abi.MMHyper) else abi.MMHyper(*params)
    strain, N = _n(strain, layout)
    S, dSdC = _alloc(6, N, layout), _alloc(36, N, layout)
    rc = lib().mmo_hyper(model, C.byref(p), _ptr(strain), strain_kind, _ptr(S), _ptr(dSdC), N, layout)
    assert rc == 0
    return S, dSdC


def hyper_ad(model, params, C6):
    p = params if isinstance(params, abi.MMHyper) else abi.MMHyper(*params)
    C6 = np.ascontiguousarray(C6, dtype=np.float64)
    psi, S, T = np.zeros(1), np.zeros(6), np.zeros(36)
    lib().mmo_hyper_ad(model, C.byref(p), _ptr(C6), _ptr(psi), _ptr(S), _ptr(T))
    return float(psi[0]), S, T


def strain_energy(model, params, C6):
    p = params if isinstance(params, abi.MMHyper) else abi.MMHyper(*params)
    C6 = np.ascontiguousarray(C6, dtype=np.float64)
    return lib().mmo_strain_energy(model, C.byref(p), _ptr(C6))


def crystal(p, deps, state, dt=1.0, layout=SOA, tol=1e-8, max_iter=100):
    deps, N = _n(deps, layout)
    state = np.ascontiguousarray(state, dtype=np.float64)
    sig, tan, st = _alloc(6, N, layout), _alloc(36, N, layout), _alloc(42, N, layout)
    active, iters = np.zeros(N, np.uint16), np.zeros(N, np.uint16)
    nfail = np.zeros(1, np.int32)
    o = abi.MMNewtonOpts(tol, max_iter, 0)
    rc = lib().mmo_crystal(C.byref(p), _ptr(deps), dt, _ptr(state), _ptr(sig), _ptr(tan), _ptr(st), _ptr(active), _ptr(iters), _ptr(nfail), C.byref(o), N, layout)
    assert rc == 0
    return dict(sig=sig, tan=tan, state=st, active=active, iters=iters, n_fail=int(nfail[0]))


def xu_needleman(p, jump, dim=3, layout=SOA):
    jump, N = _n(jump, layout)
    T, H = _alloc(dim, N, layout), _alloc(dim * dim, N, layout)
    rc = lib().mmo_xu_needleman(C.byref(p), _ptr(jump), _ptr(T), _ptr(H), dim, N, layout)
    assert rc == 0
    return T, H


def solve_scalar_kat(which, x0):
    x, d = np.zeros(1), np.zeros(1)
    it = np.zeros(1, np.int32)
    rc = lib().mmo_solve_scalar_kat(which, x0, _ptr(x), _ptr(d), _ptr(it))
    return rc, float(x[0]), float(d[0]), int(it[0])


def synth_uniform(N, nc, seed, lo, hi, layout=SOA):
    out = _alloc(nc, N, layout)
    lo = np.ascontiguousarray(np.broadcast_to(np.asarray(lo, dtype=np.float64), (nc,)))
    hi = np.ascontiguousarray(np.broadcast_to(np.asarray(hi, dtype=np.float64), (nc,)))
    lib().mmo_synth_uniform(_ptr(out), N, nc, layout, seed, _ptr(lo), _ptr(hi))
    return out


# ---- dimension wrappers (reference src/wrappers.jl) ---------------------------------------------
UNIAXIAL_STRAIN, PLANE_STRAIN, UNIAXIAL_STRESS, PLANE_STRESS, SHELL_STRESS = 1, 2, 3, 4, 5
WRAP_NCOMP = {1: 1, 2: 3, 3: 1, 4: 3, 5: 6}


def wrapped_linear_elastic(kind, E, nu, eps_red, layout=SOA, tol=1e-8, max_iter=10):
    nc = WRAP_NCOMP[kind]
    eps_red, N = _n(eps_red, layout)
    sig, tan = _alloc(nc, N, layout), _alloc(nc * nc, N, layout)
    oit, nfail = np.zeros(N, np.uint8), np.zeros(1, np.int32)
    p = abi.MMLinearElastic(E, nu)
    w = abi.MMNewtonOpts(tol, max_iter, 0)
    rc = lib().mmo_wrapped_linear_elastic(kind, C.byref(p), _ptr(eps_red), _ptr(sig), _ptr(tan), _ptr(oit), _ptr(nfail), C.byref(w), N, layout)
    assert rc == 0
    return dict(sig=sig, tan=tan, outer_iters=oit, n_fail=int(nfail[0]))


def wrapped_plastic(kind, params, eps_red, state, layout=SOA, tol=1e-8, max_iter=1000, wtol=1e-8, wmax_iter=10):
    nc = WRAP_NCOMP[kind]
    p = params if isinstance(params, abi.MMPlastic) else abi.MMPlastic(*params)
    eps_red, N = _n(eps_red, layout)
    state = np.ascontiguousarray(state, dtype=np.float64)
    sig, tan, st = _alloc(nc, N, layout), _alloc(nc * nc, N, layout), _alloc(14, N, layout)
    flags, iters, oit = np.zeros(N, np.uint8), np.zeros(N, np.uint16), np.zeros(N, np.uint8)
    nfail = np.zeros(1, np.int32)
    o, w = abi.MMNewtonOpts(tol, max_iter, 0), abi.MMNewtonOpts(wtol, wmax_iter, 0)
    rc = lib().mmo_wrapped_plastic(kind, C.byref(p), _ptr(eps_red), _ptr(state), _ptr(sig), _ptr(tan), _ptr(st), _ptr(flags), _ptr(iters),
                                   _ptr(oit), _ptr(nfail), C.byref(o), C.byref(w), N, layout)
    assert rc == 0
    return dict(sig=sig, tan=tan, state=st, flags=flags, iters=iters, outer_iters=oit, n_fail=int(nfail[0]))


# ---- trait layer (reference src/traits.jl) -------------------------------------------------------
TAN_DSDC, TAN_DSDE, TAN_DPDF = 0, 1, 2
STRAIN_C, STRAIN_F, STRAIN_E = 0, 1, 2


def hyper_ex(model, params, strain, strain_kind, tangent_kind, layout=SOA):
    p = params if isinstance(params, abi.MMHyper) else abi.MMHyper(*params)
    strain, N = _n(strain, layout)
    ns, nt = (9, 81) if tangent_kind == TAN_DPDF else (6, 36)
    S, T = _alloc(ns, N, layout), _alloc(nt, N, layout)
    rc = lib().mmo_hyper_ex(model, C.byref(p), _ptr(strain), strain_kind, tangent_kind, _ptr(S), _ptr(T), N, layout)
    assert rc == 0
    return S, T


def neohook_dPdF_ad(params, F9):
    p = params if isinstance(params, abi.MMHyper) else abi.MMHyper(*params)
    F9 = np.ascontiguousarray(F9, dtype=np.float64)
    Pt, dP = np.zeros(9), np.zeros(81)
    lib().mmo_neohook_dPdF_ad(C.byref(p), _ptr(F9), _ptr(Pt), _ptr(dP))
    return Pt, dP


# ---- TransverselyIsotropic and energies ----------------------------------
def transviso_params(E_L, E_T, G_LT, nu_LT, nu_TT):
    p = abi.MMTransverselyIsotropic()
    lib().mmo_transviso_setup(E_L, E_T, G_LT, nu_LT, nu_TT, C.byref(p))
    return p


def transversely_isotropic(p, eps, a3, layout=SOA):
    eps, N = _n(eps, layout)
    a3 = np.ascontiguousarray(a3, dtype=np.float64)
    sig, tan = _alloc(6, N, layout), _alloc(36, N, layout)
    rc = lib().mmo_transversely_isotropic(C.byref(p), _ptr(eps), _ptr(a3), _ptr(sig), _ptr(tan), N, layout)
    assert rc == 0
    return sig, tan


def strain_energy_batch(model, params, strain, strain_kind=0, layout=SOA):
    """model = -1: LinearElastic (params = (E, nu)); else hyper model with strain kind C/F/E"""
    strain, N = _n(strain, layout)
    psi = np.zeros(N)
    p = (abi.MMLinearElastic(*params) if model < 0 else (params if isinstance(params, abi.MMHyper) else abi.MMHyper(*params)))
    rc = lib().mmo_strain_energy_batch(model, C.byref(p), _ptr(strain), strain_kind, _ptr(psi), N, layout)
    assert rc == 0
    return psi


def wrapped_transviso(kind, p, eps_red, a3, layout=SOA, tol=1e-8, max_iter=10):
    nc = WRAP_NCOMP[kind]
    eps_red, N = _n(eps_red, layout)
    a3 = np.ascontiguousarray(a3, dtype=np.float64)
    sig, tan = _alloc(nc, N, layout), _alloc(nc * nc, N, layout)
    oit, nfail = np.zeros(N, np.uint8), np.zeros(1, np.int32)
    w = abi.MMNewtonOpts(tol, max_iter, 0)
    rc = lib().mmo_wrapped_transviso(kind, C.byref(p), _ptr(eps_red), _ptr(a3), _ptr(sig), _ptr(tan), _ptr(oit), _ptr(nfail), C.byref(w), N, layout)
    assert rc == 0
    return dict(sig=sig, tan=tan, outer_iters=oit, n_fail=int(nfail[0]))


def cells(model, params, nn, nq, dNdx, dOmega, ue, state=None, want_tangent=True, want_stiffness=False, tol=1e-8, max_iter=1000):
    """the caller's quadrature loop around material_response.  model 0: LinearElastic (E, nu); 1: Plastic.
    dNdx (3nn, P), dOmega (P,), ue (3nn, ncell), state (14, P) with P = ncell*nq."""
    dNdx = np.ascontiguousarray(dNdx, dtype=np.float64); dOmega = np.ascontiguousarray(dOmega, dtype=np.float64); ue = np.ascontiguousarray(ue, dtype=np.float64)
    ncell = ue.shape[1]; P = ncell * nq
    assert dNdx.shape == (3 * nn, P) and dOmega.shape == (P,) and ue.shape == (3 * nn, ncell)
    fe, sig = np.zeros((3 * nn, ncell)), np.zeros((6, P))
    tan = np.zeros((36, P)) if want_tangent else None
    ke = np.zeros((9 * nn * nn, ncell)) if want_stiffness else None
    if model == 0:
        p = abi.MMLinearElastic(*params)
        st_in = st_out = flags = iters = None
    elif model == 2:
        p = params
        st_in = np.ascontiguousarray(state, dtype=np.float64)          # a3 (3, P)
        st_out = flags = iters = None
    else:
        p = params if isinstance(params, abi.MMPlastic) else abi.MMPlastic(*params)
        st_in = np.ascontiguousarray(state, dtype=np.float64); st_out = np.zeros((14, P))
        flags, iters = np.zeros(P, np.uint8), np.zeros(P, np.uint16)
    nfail = np.zeros(1, np.int32)
    o = abi.MMNewtonOpts(tol, max_iter, 0)
    rc = lib().mmo_cells(model, C.byref(p), nn, nq, _ptr(dNdx), _ptr(dOmega), _ptr(ue), _ptr(st_in), _ptr(fe), _ptr(ke), _ptr(sig), _ptr(tan), _ptr(st_out),
                         _ptr(flags), _ptr(iters), _ptr(nfail), C.byref(o), ncell)
    assert rc == 0
    return dict(fe=fe, ke=ke, sig=sig, tan=tan, state=st_out, flags=flags, iters=iters, n_fail=int(nfail[0]))


def cells_hyper(model, params, nn, nq, dNdx, dOmega, ue, want_tangent=True):
    """total-Lagrangian cell loop for NeoHook / Yeoh / StVenant: F = I + Σ u ⊗ ∇N, (Pᵀ, ∂Pᵀ∂F), fe = Σ P⋅∇N dΩ"""
    dNdx = np.ascontiguousarray(dNdx, dtype=np.float64); dOmega = np.ascontiguousarray(dOmega, dtype=np.float64); ue = np.ascontiguousarray(ue, dtype=np.float64)
    ncell = ue.shape[1]; P = ncell * nq
    assert dNdx.shape == (3 * nn, P) and dOmega.shape == (P,) and ue.shape == (3 * nn, ncell)
    p = params if isinstance(params, abi.MMHyper) else abi.MMHyper(*params)
    fe, Pt = np.zeros((3 * nn, ncell)), np.zeros((9, P))
    dP = np.zeros((81, P)) if want_tangent else None
    rc = lib().mmo_cells_hyper(model, C.byref(p), nn, nq, _ptr(dNdx), _ptr(dOmega), _ptr(ue), _ptr(fe), _ptr(Pt), _ptr(dP), ncell)
    assert rc == 0
    return dict(fe=fe, Pt=Pt, dPdF=dP)


def wrapped_crystal(kind, p, deps_red, state, dt=1.0, layout=SOA, tol=1e-8, max_iter=100, wtol=1e-8, wmax_iter=10):
    nc = WRAP_NCOMP[kind]
    deps_red, N = _n(deps_red, layout)
    state = np.ascontiguousarray(state, dtype=np.float64)
    sig, tan, st = _alloc(nc, N, layout), _alloc(nc * nc, N, layout), _alloc(42, N, layout)
    active, iters, oit = np.zeros(N, np.uint16), np.zeros(N, np.uint16), np.zeros(N, np.uint8)
    nfail = np.zeros(1, np.int32)
    o, w = abi.MMNewtonOpts(tol, max_iter, 0), abi.MMNewtonOpts(wtol, wmax_iter, 0)
    rc = lib().mmo_wrapped_crystal(kind, C.byref(p), _ptr(deps_red), dt, _ptr(state), _ptr(sig), _ptr(tan), _ptr(st), _ptr(active), _ptr(iters), _ptr(oit),
                                   _ptr(nfail), C.byref(o), C.byref(w), N, layout)
    assert rc == 0
    return dict(sig=sig, tan=tan, state=st, active=active, iters=iters, outer_iters=oit, n_fail=int(nfail[0]))
