"""B200-native batched `material_response` (MaterialModels.jl hot path) — host-side mirror.

Importing the package never touches CUDA; the native library (csrc/ -> libmmb200.so) is loaded on
first use and the load FAILS LOUDLY if it is missing — there is no CPU fallback in the product.
"""
from . import _abi, synth  # noqa: F401
from .materials import (  # noqa: F401
    AbstractDim, AbstractMaterial, AbstractMaterialState, AbstractTangent, BCC, Dim, FCC, PlaneStrain, PlaneStress, ShellStress, UniaxialStrain, UniaxialStress, CrystalViscoPlastic, DeformationGradient,
    GreenLagrange, LinearElastic, NeoHook, Plastic, RightCauchyGreen, SmallStrain, StVenant, StrainMeasure, TransverselyIsotropic, XuNeedleman,
    Yeoh, dPdF, dSdC, dSdE, dsigmadeps, native_strain_type, native_stress_type, native_tangent_type, slipsystems,
    xu_needleman_Φₙ, xu_needleman_Φₜ, xu_needleman_σₘₐₓ, xu_needleman_τₘₐₓ,
)
from .api import (  # noqa: F401
    NOT_CONVERGED, PLANE_STRESS_NOT_CONVERGED, CompactPlasticState, CompactTangent, CrystalViscoPlasticState, PinnedArray, LinearElasticState, NeoHookState, PlasticCache, PlasticState, StVenantState,
    TransverselyIsotropicState, XuNeedlemanState, YeohState, cell_internal_forces, elastic_strain_energy_density, elastic_tangent_3D, get_cache, initial_material_state, material_response, synth_uniform,
)

__version__ = "0.1.0"
