"""Material types, states and strain/tangent tags — the host-side mirror of the reference's Julia
types (same names, same keyword constructors, same field meaning).  Unicode keyword names of the
reference (ν, σ_y, κ_∞, ...) are accepted next to ASCII aliases.
"""
import ctypes
import math
import unicodedata

from . import _abi


def _kw(kwargs, *names, default=None, required=True):
    # Python NFKC-normalises identifiers (σₘₐₓ= arrives as "σmax"), dict-splatted keys do not: accept both
    norm = {unicodedata.normalize("NFKC", k): k for k in kwargs}
    for n in names:
        k = norm.get(unicodedata.normalize("NFKC", n))
        if k is not None:
            return float(kwargs.pop(k))
    if required and default is None:
        raise TypeError("missing keyword argument %s" % names[0])
    return default


class AbstractMaterial:
    pass


class AbstractMaterialState:
    pass


# --------------------------------------------------------------------------------------------
class LinearElastic(AbstractMaterial):
    """LinearElastic(E, ν) — reference src/LinearElastic.jl:10-22."""

    def __init__(self, *args, **kw):
        if args:
            self.E, self.ν = map(float, args)
        else:
            self.E = _kw(kw, "E")
            self.ν = _kw(kw, "ν", "nu")
        self.nu = self.ν
        self._c = _abi.MMLinearElastic(self.E, self.ν)

    @property
    def Eᵉ(self):
        from .api import elastic_tangent_3D

        return elastic_tangent_3D(self.E, self.ν)


class Plastic(AbstractMaterial):
    """Plastic(E, ν, σ_y, H, r, κ_∞, α_∞) — reference src/Plastic.jl:15-39."""

    def __init__(self, *args, **kw):
        if args:
            self.E, self.ν, self.σ_y, self.H, self.r, self.κ_inf, self.α_inf = map(float, args)
        else:
            self.E = _kw(kw, "E")
            self.ν = _kw(kw, "ν", "nu")
            self.σ_y = _kw(kw, "σ_y", "sigma_y")
            self.H = _kw(kw, "H")
            self.r = _kw(kw, "r")
            self.κ_inf = _kw(kw, "κ_∞", "κ_inf", "kappa_inf")
            self.α_inf = _kw(kw, "α_∞", "α_inf", "alpha_inf")
        self.H_κ = -self.r * self.H                    # Plastic.jl:31
        self.H_α = -2 / 3 * (1 - self.r) * self.H      # Plastic.jl:32
        self._c = _abi.MMPlastic(self.E, self.ν, self.σ_y, self.H, self.r, self.κ_inf, self.α_inf)

    @property
    def Eᵉ(self):
        """the one stored elastic tangent every elastic point returns — reference src/Plastic.jl:23, :135"""
        from .api import elastic_tangent_3D

        return elastic_tangent_3D(self.E, self.ν)


class _Hyper(AbstractMaterial):
    MODEL = None

    def __init__(self, *args, **kw):
        if args:
            vals = list(map(float, args)) + [0.0, 0.0]
            self.λ, self.μ, self.c2, self.c3 = vals[:4]
        else:
            self.λ = _kw(kw, "λ", "lambda_", "lam")
            self.μ = _kw(kw, "μ", "mu")
            self.c2 = _kw(kw, "c₂", "c2", default=0.0, required=False) or 0.0
            self.c3 = _kw(kw, "c₃", "c3", default=0.0, required=False) or 0.0
        self._c = _abi.MMHyper(self.λ, self.μ, self.c2, self.c3)


class NeoHook(_Hyper):
    """NeoHook(; λ, μ) — reference src/FiniteStrain/neohook.jl:10-25."""
    MODEL = _abi.MM_NEOHOOK


class Yeoh(_Hyper):
    """Yeoh(; λ, μ, c₂, c₃) — reference src/FiniteStrain/yeoh.jl:11-27."""
    MODEL = _abi.MM_YEOH


class StVenant(_Hyper):
    """StVenant(; λ, μ) — reference src/FiniteStrain/stvenant.jl:10-24."""
    MODEL = _abi.MM_STVENANT


class TransverselyIsotropic(AbstractMaterial):
    """TransverselyIsotropic(; E_L, E_T, G_LT, ν_LT, ν_TT) or TransverselyIsotropic(L⊥, L₌, M₌, G⊥, G₌)
    — reference src/transverselyisotropic.jl:19-53.  The material direction lives in the state."""

    def __init__(self, *args, **kw):
        if args:
            if len(args) != 5:
                raise TypeError("TransverselyIsotropic(L⊥, L₌, M₌, G⊥, G₌) takes 5 positional arguments")
            c = _abi.MMTransverselyIsotropic(*map(float, args))
        else:
            from ._lib import check, load

            E_L, E_T, G_LT = _kw(kw, "E_L"), _kw(kw, "E_T"), _kw(kw, "G_LT")
            ν_LT, ν_TT = _kw(kw, "ν_LT", "nu_LT"), _kw(kw, "ν_TT", "nu_TT")
            c = _abi.MMTransverselyIsotropic()
            check(load().mm_transversely_isotropic_setup(E_L, E_T, G_LT, ν_LT, ν_TT, ctypes.byref(c)), "mm_transversely_isotropic_setup")
        self._c = c
        self.L_perp, self.L_par, self.M_par, self.G_perp, self.G_par = c.L_perp, c.L_par, c.M_par, c.G_perp, c.G_par


class FCC:
    CODE = _abi.MM_FCC


class BCC:
    CODE = _abi.MM_BCC


class CrystalViscoPlastic(AbstractMaterial):
    """CrystalViscoPlastic(E, ν, τ_y, H_iso, H_kin, q, α_∞, t_star, σ_c, m, slipsystems)
    — reference src/CrystalViscoPlastic/CrystalViscoPlastic.jl:18-48.  `slipsystems` is the list of
    (plane normal, direction) pairs returned by `slipsystems(FCC()/BCC(), rodrigues)`; S must be 12."""

    def __init__(self, **kw):
        from ._lib import load, check

        self.E = _kw(kw, "E")
        self.ν = _kw(kw, "ν", "nu")
        self.τ_y = _kw(kw, "τ_y", "tau_y")
        self.H_iso = _kw(kw, "H_iso")
        self.H_kin = _kw(kw, "H_kin")
        self.q = _kw(kw, "q")
        self.α_inf = _kw(kw, "α_∞", "α_inf", "alpha_inf")
        self.t_star = _kw(kw, "t_star")
        self.σ_c = _kw(kw, "σ_c", "sigma_c")
        self.m = _kw(kw, "m")
        self.slipsystems = kw.pop("slipsystems")
        self.nslipsystems = len(self.slipsystems)
        if self.nslipsystems != _abi.MM_NSLIP:
            raise ValueError("CrystalViscoPlastic: this build supports S = %d slip systems" % _abi.MM_NSLIP)
        c = _abi.MMCrystal(E=self.E, nu=self.ν, tau_y=self.τ_y, H_iso=self.H_iso, H_kin=self.H_kin, q=self.q,
                           alpha_inf=self.α_inf, t_star=self.t_star, sigma_c=self.σ_c, m=self.m)
        planes = (ctypes.c_double * 36)(*[x for ms in self.slipsystems for x in ms[0]])
        dirs = (ctypes.c_double * 36)(*[x for ms in self.slipsystems for x in ms[1]])
        check(load().mm_crystal_setup(ctypes.byref(c), planes, dirs), "mm_crystal_setup")
        self._c = c


def slipsystems(structure, rodrigues):
    """slipsystems(::FCC/::BCC, R) — reference slipsystems.jl:37-53; R given as RodriguesParam (x,y,z)."""
    from ._lib import load

    code = structure.CODE if hasattr(structure, "CODE") else int(structure)
    r = (ctypes.c_double * 3)(*map(float, rodrigues))
    planes = (ctypes.c_double * 36)()
    dirs = (ctypes.c_double * 36)()
    n = load().mm_slipsystems(code, r, planes, dirs)
    if n != _abi.MM_NSLIP:
        raise RuntimeError("mm_slipsystems failed")
    return [(tuple(planes[3 * s:3 * s + 3]), tuple(dirs[3 * s:3 * s + 3])) for s in range(n)]


# --- XuNeedleman parameter conversions (XuNeedleman.jl:42-60) ---
def xu_needleman_Φₙ(σₘₐₓ, δₙ):
    return σₘₐₓ * math.exp(1.0) * δₙ


def xu_needleman_Φₜ(τₘₐₓ, δₜ):
    return τₘₐₓ * math.sqrt(0.5 * math.exp(1.0)) * δₜ


def xu_needleman_σₘₐₓ(Φₙ, δₙ):
    return Φₙ / (math.exp(1.0) * δₙ)


def xu_needleman_τₘₐₓ(Φₜ, δₜ):
    return Φₜ / (math.sqrt(0.5 * math.exp(1.0)) * δₜ)


class XuNeedleman(AbstractMaterial):
    """XuNeedleman(; σₘₐₓ, τₘₐₓ, Φₙ, Φₜ, Δₙˢ) — reference src/CohesiveMaterials/XuNeedleman.jl:22-34."""

    def __init__(self, **kw):
        σmax = _kw(kw, "σₘₐₓ", "sigma_max")
        τmax = _kw(kw, "τₘₐₓ", "tau_max")
        self.Φₙ = _kw(kw, "Φₙ", "Phi_n")
        self.Φₜ = _kw(kw, "Φₜ", "Phi_t")
        self.Δₙˢ = _kw(kw, "Δₙˢ", "Delta_n_star")
        self.δₙ = self.Φₙ / (math.exp(1.0) * σmax)
        self.δₜ = self.Φₜ / (math.sqrt(0.5 * math.exp(1.0)) * τmax)
        self._c = _abi.MMXuNeedleman(self.Φₙ, self.Φₜ, self.δₙ, self.δₜ, self.Δₙˢ)


# --------------------------------------------------------------------------------------------
# dimension wrappers (reference src/wrappers.jl:1-15)
class AbstractDim:
    KIND = 0
    DIM = 3


class UniaxialStrain(AbstractDim):
    KIND, DIM = _abi.MM_DIM_UNIAXIAL_STRAIN, 1


class UniaxialStress(AbstractDim):
    KIND, DIM = _abi.MM_DIM_UNIAXIAL_STRESS, 1


class PlaneStrain(AbstractDim):
    KIND, DIM = _abi.MM_DIM_PLANE_STRAIN, 2


class PlaneStress(AbstractDim):
    KIND, DIM = _abi.MM_DIM_PLANE_STRESS, 2


class ShellStress(AbstractDim):
    KIND, DIM = _abi.MM_DIM_SHELL_STRESS, 3


class Dim(AbstractDim):
    """Dim{d}: the generic strain-restricted state — reference src/wrappers.jl:3, dispatched together with UniaxialStrain / PlaneStrain
    at :29-41 (pad the d-dimensional strain with zeros, cut the 3D result down).  Dim(1) == UniaxialStrain, Dim(2) == PlaneStrain storage;
    Dim(3) is the plain 3D call (reduce_dim(A, ::AbstractDim{3}) = A, :22)."""

    def __init__(self, d):
        d = int(d)
        if d not in (1, 2, 3):
            raise ValueError("Dim(d): d must be 1, 2 or 3")
        self.DIM = d
        self.KIND = {1: _abi.MM_DIM_UNIAXIAL_STRAIN, 2: _abi.MM_DIM_PLANE_STRAIN, 3: 0}[d]


# --------------------------------------------------------------------------------------------
# strain measures and tangent tags (reference src/traits.jl:1-26)
class StrainMeasure:
    def __init__(self, value):
        self.value = value


class DeformationGradient(StrainMeasure):
    pass


class RightCauchyGreen(StrainMeasure):
    pass


class GreenLagrange(StrainMeasure):
    pass


class SmallStrain(StrainMeasure):
    pass


class AbstractTangent:
    pass


class dSdC(AbstractTangent):       # ∂S∂C
    pass


class dSdE(AbstractTangent):       # ∂S∂E
    pass


class dPdF(AbstractTangent):       # ∂Pᵀ∂F
    pass


class dsigmadeps(AbstractTangent):  # ∂σ∂ε
    pass


def native_tangent_type(M):
    """native_tangent_type(::Type{M}) — reference neohook.jl:27, yeoh.jl:29, stvenant.jl:26 (∂S∂C), LinearElastic.jl:25 (∂σ∂ε).  The
    reference defines it for LinearElastic only among the small-strain models; here every small-strain model answers ∂σ∂ε."""
    if issubclass(M, _Hyper):
        return dSdC
    return dsigmadeps


def native_strain_type(M):
    """native_strain_type(::Type{M}) — reference src/traits.jl:138,140-143 (test_linear_elastic.jl:32)"""
    return RightCauchyGreen if native_tangent_type(M) is dSdC else SmallStrain


def native_stress_type(M):
    """native_stress_type(::Type{M}) — reference src/traits.jl:139-143: the NAME of the stress measure (no wrapper types are needed here)"""
    return "SecondPiolaKirchhoff" if native_tangent_type(M) is dSdC else "TrueStress"
