"""B200-native batched `material_response` (MaterialModels.jl hot path) — host-side mirror.

Importing the package never touches CUDA; the native library (csrc/ -> libmmb200.so) is loaded on
first use and the load FAILS LOUDLY if it is missing — there is no CPU fallback in the product.
"""
from . import _abi  # noqa: F401
