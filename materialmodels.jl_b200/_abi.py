"""ctypes mirror of include/mmb200.h (POD parameter structs + constants).  Pure ctypes: importable
without a GPU and without the CUDA library (the oracle's Python binding reuses these structs)."""
import ctypes as C

MM_LAYOUT_SOA = 0
MM_LAYOUT_AOS = 1
MM_NSLIP = 12
MM_OK, MM_ERR_ARG, MM_ERR_CUDA, MM_ERR_NOGPU, MM_ERR_CAPACITY = 0, -1, -2, -3, -4
MM_ROWS_TANGENT_STATE, MM_ROWS_FACTORED_STATE, MM_ROWS_TANGENT, MM_ROWS_FACTORED = 0, 1, 2, 3
MM_NEOHOOK, MM_YEOH, MM_STVENANT = 0, 1, 2
MM_STRAIN_C, MM_STRAIN_F, MM_STRAIN_E = 0, 1, 2
MM_TANGENT_DSDC, MM_TANGENT_DSDE, MM_TANGENT_DPDF = 0, 1, 2
MM_FCC, MM_BCC = 0, 1
MM_FLAG_PLASTIC = 0x01
MM_FLAG_FAILED = 0x80
MM_ACTIVE_FAILED = 0x8000
MM_FLAG_WRAPPER_FAILED = 0x40
MM_ACTIVE_WRAPPER_FAILED = 0x2000
MM_DIM_UNIAXIAL_STRAIN, MM_DIM_PLANE_STRAIN, MM_DIM_UNIAXIAL_STRESS, MM_DIM_PLANE_STRESS, MM_DIM_SHELL_STRESS = 1, 2, 3, 4, 5


class MMLinearElastic(C.Structure):
    _fields_ = [("E", C.c_double), ("nu", C.c_double)]


class MMPlastic(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("E", "nu", "sigma_y", "H", "r", "kappa_inf", "alpha_inf")]


class MMHyper(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("lambda_", "mu", "c2", "c3")]


class MMXuNeedleman(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("Phi_n", "Phi_t", "delta_n", "delta_t", "Delta_n_star")]


class MMCrystal(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("E", "nu", "tau_y", "H_iso", "H_kin", "q", "alpha_inf", "t_star", "sigma_c", "m")] + [
        ("nslip", C.c_int32),
        ("pad_", C.c_int32),
        ("MS", (C.c_double * 9) * MM_NSLIP),
        ("H", (C.c_double * MM_NSLIP) * MM_NSLIP),
    ]


class MMTransverselyIsotropic(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("L_perp", "L_par", "M_par", "G_perp", "G_par")]


class MMNewtonOpts(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_iter", C.c_int32), ("pad_", C.c_int32)]


# name -> (restype, argtypes) for every symbol include/mmb200.h declares
_P = C.c_void_p
_I64 = C.c_int64
SIGNATURES = {
    "mm_version": (C.c_char_p, []),
    "mm_last_error": (C.c_char_p, []),
    "mm_device_count": (C.c_int, []),
    "mm_set_tuning": (C.c_int, [C.c_char_p, C.c_int]),
    "mm_elastic_tangent_3d": (C.c_int, [C.c_double, C.c_double, _P]),
    "mm_slipsystems": (C.c_int, [C.c_int, _P, _P, _P]),
    "mm_crystal_setup": (C.c_int, [C.POINTER(MMCrystal), _P, _P]),
    "mm_xu_needleman_setup": (C.c_int, [C.c_double] * 5 + [C.POINTER(MMXuNeedleman)]),
    "mm_linear_elastic": (C.c_int, [C.POINTER(MMLinearElastic), _P, _P, _P, _I64, C.c_int, _P]),
    "mm_plastic": (C.c_int, [C.POINTER(MMPlastic), _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(MMNewtonOpts), _I64, C.c_int, _P]),
    "mm_hyper": (C.c_int, [C.c_int, C.POINTER(MMHyper), _P, C.c_int, _P, _P, _I64, C.c_int, _P]),
    "mm_hyper_ex": (C.c_int, [C.c_int, C.POINTER(MMHyper), _P, C.c_int, C.c_int, _P, _P, _I64, C.c_int, _P]),
    "mm_crystal": (C.c_int, [C.POINTER(MMCrystal), _P, C.c_double, _P, _P, _P, _P, _P, _P, _P, C.POINTER(MMNewtonOpts), _I64, C.c_int, _P]),
    "mm_xu_needleman": (C.c_int, [C.POINTER(MMXuNeedleman), _P, _P, _P, C.c_int, _I64, C.c_int, _P]),
    "mm_wrap_ncomp": (C.c_int, [C.c_int]),
    "mm_wrapped_linear_elastic": (C.c_int, [C.c_int, C.POINTER(MMLinearElastic), _P, _P, _P, _P, _P, C.POINTER(MMNewtonOpts), _I64, C.c_int, _P]),
    "mm_wrapped_plastic": (C.c_int, [C.c_int, C.POINTER(MMPlastic), _P, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(MMNewtonOpts), C.POINTER(MMNewtonOpts),
                                     _I64, C.c_int, _P]),
    "mm_wrapped_crystal": (C.c_int, [C.c_int, C.POINTER(MMCrystal), _P, C.c_double, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(MMNewtonOpts),
                                     C.POINTER(MMNewtonOpts), _I64, C.c_int, _P]),
    "mm_wrapped_transversely_isotropic": (C.c_int, [C.c_int, C.POINTER(MMTransverselyIsotropic), _P, _P, _P, _P, _P, _P, C.POINTER(MMNewtonOpts), _I64, C.c_int, _P]),
    "mmh_wrapped_linear_elastic": (C.c_int, [C.c_int, C.POINTER(MMLinearElastic), _P, _P, _P, _P, _P, C.POINTER(MMNewtonOpts), _I64, C.c_int]),
    "mmh_wrapped_plastic": (C.c_int, [C.c_int, C.POINTER(MMPlastic), _P, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(MMNewtonOpts), C.POINTER(MMNewtonOpts),
                                      _I64, C.c_int]),
    "mmh_wrapped_crystal": (C.c_int, [C.c_int, C.POINTER(MMCrystal), _P, C.c_double, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(MMNewtonOpts),
                                      C.POINTER(MMNewtonOpts), _I64, C.c_int]),
    "mmh_wrapped_transversely_isotropic": (C.c_int, [C.c_int, C.POINTER(MMTransverselyIsotropic), _P, _P, _P, _P, _P, _P, C.POINTER(MMNewtonOpts), _I64, C.c_int]),
    "mm_transversely_isotropic_setup": (C.c_int, [C.c_double] * 5 + [C.POINTER(MMTransverselyIsotropic)]),
    "mm_transversely_isotropic": (C.c_int, [C.POINTER(MMTransverselyIsotropic), _P, _P, _P, _P, _I64, C.c_int, _P]),
    "mm_linear_elastic_energy": (C.c_int, [C.POINTER(MMLinearElastic), _P, _P, _I64, C.c_int, _P]),
    "mm_hyper_energy": (C.c_int, [C.c_int, C.POINTER(MMHyper), _P, C.c_int, _P, _I64, C.c_int, _P]),
    "mm_cells_linear_elastic": (C.c_int, [C.POINTER(MMLinearElastic), C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "mm_cells_plastic": (C.c_int, [C.POINTER(MMPlastic), C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(MMNewtonOpts), _I64, _P]),
    "mm_cells_transversely_isotropic": (C.c_int, [C.POINTER(MMTransverselyIsotropic), C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "mm_cells_hyper": (C.c_int, [C.c_int, C.POINTER(MMHyper), C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "mm_synth_uniform": (C.c_int, [_P, _I64, C.c_int, C.c_int, C.c_uint64, _P, _P, _P]),
    "mmh_hyper_ex": (C.c_int, [C.c_int, C.POINTER(MMHyper), _P, C.c_int, C.c_int, _P, _P, _I64, C.c_int]),
    "mmh_transversely_isotropic": (C.c_int, [C.POINTER(MMTransverselyIsotropic), _P, _P, _P, _P, _I64, C.c_int]),
    "mmh_linear_elastic_energy": (C.c_int, [C.POINTER(MMLinearElastic), _P, _P, _I64, C.c_int]),
    "mmh_hyper_energy": (C.c_int, [C.c_int, C.POINTER(MMHyper), _P, C.c_int, _P, _I64, C.c_int]),
    "mmh_alloc_pinned": (_P, [C.c_uint64]),
    "mmh_free_pinned": (None, [_P]),
    "mmh_linear_elastic": (C.c_int, [C.POINTER(MMLinearElastic), _P, _P, _P, _I64, C.c_int]),
    "mmh_plastic": (C.c_int, [C.POINTER(MMPlastic), _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(MMNewtonOpts), _I64, C.c_int]),
    "mmh_hyper": (C.c_int, [C.c_int, C.POINTER(MMHyper), _P, C.c_int, _P, _P, _I64, C.c_int]),
    "mmh_crystal": (C.c_int, [C.POINTER(MMCrystal), _P, C.c_double, _P, _P, _P, _P, _P, _P, _P, C.POINTER(MMNewtonOpts), _I64, C.c_int]),
    "mmh_crystal_resident": (C.c_int, [C.POINTER(MMCrystal), _P, C.c_double, _P, _P, _P, _P, _P, _P, C.POINTER(MMNewtonOpts), _I64, C.c_int]),
    "mmh_xu_needleman": (C.c_int, [C.POINTER(MMXuNeedleman), _P, _P, _P, C.c_int, _I64, C.c_int]),
    "mmh_set_chunk_points": (_I64, [_I64]),
    "mm_plastic_row_width": (C.c_int, [C.c_int]),
    "mm_plastic_compact_workspace": (C.c_uint64, [_I64, C.c_int]),
    "mm_plastic_compact": (C.c_int, [C.POINTER(MMPlastic), _P, _P, _P, _P, _P, C.c_int, _P, _I64, _P, _P, _P, C.POINTER(MMNewtonOpts), _P, C.c_uint64, _I64, C.c_int, _P]),
    "mmh_plastic_compact": (C.c_int, [C.POINTER(MMPlastic), _P, _P, _P, _P, _P, C.c_int, _P, _I64, _P, _P, _P, C.POINTER(MMNewtonOpts), _I64, C.c_int]),
    "mmh_plastic_compact_resident": (C.c_int, [C.POINTER(MMPlastic), _P, _P, _P, _P, _P, C.c_int, _P, _I64, _P, _P, C.POINTER(MMNewtonOpts), _I64, C.c_int]),
    "mmh_plastic_expand": (C.c_int, [C.POINTER(MMPlastic), C.c_int, _P, _P, _I64, _P, _P, _P, _I64, C.c_int]),
    "mmh_set_host_threads": (C.c_int, [C.c_int]),
}


def bind(lib, signatures=SIGNATURES, prefix_from="mm", prefix_to=None):
    """Attach restype/argtypes; raises AttributeError if a declared symbol is missing."""
    for name, (res, args) in signatures.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
