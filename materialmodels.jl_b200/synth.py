"""Deterministic synthetic workloads: the inputs of BASELINE.json's five configs.

u(i,c) = (splitmix64(seed + i*nc + c) >> 11) * 2^-53, value(i,c) = lo[c] + (hi[c]-lo[c]) * u(i,c).
`host_uniform` (numpy) and the device generator `mm_synth_uniform` produce bit-identical arrays, so
the oracle and the CUDA path see the same inputs without any transfer.
"""
import numpy as np

SEED0 = 20261019
SOA, AOS = 0, 1

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return x ^ (x >> np.uint64(31))


def host_uniform(N, nc, seed, lo, hi, layout=SOA, i0=0, idx=None):
    """Points [i0, i0+N) (or the explicit point indices `idx`) of the global stream with `nc`
    components; shape (nc,N) SoA or (N,nc) AoS."""
    lo = np.broadcast_to(np.asarray(lo, dtype=np.float64), (nc,))
    hi = np.broadcast_to(np.asarray(hi, dtype=np.float64), (nc,))
    with np.errstate(over="ignore"):
        if idx is not None:
            i = np.asarray(idx).astype(np.uint64)[:, None]
        else:
            i = (np.arange(N, dtype=np.uint64) + np.uint64(i0))[:, None]
        c = np.arange(nc, dtype=np.uint64)[None, :]
        h = _splitmix64(np.uint64(seed) + i * np.uint64(nc) + c)
    u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    v = lo[None, :] + (hi - lo)[None, :] * u
    return np.ascontiguousarray(v.T) if layout == SOA else np.ascontiguousarray(v)


# ---- material parameter sets used by the reference's own tests ---------------------------------
PLASTIC_PARAMS = dict(E=200e3, nu=0.3, sigma_y=200.0, H=50.0, r=0.5, kappa_inf=13.0, alpha_inf=13.0)   # test_plastic.jl:3
ELASTIC_PARAMS = dict(E=200e3, nu=0.3)                                                                   # test_linear_elastic.jl:4-6
CRYSTAL_PARAMS = dict(E=200e3, nu=0.3, tau_y=400.0, H_iso=1e3, H_kin=1e3, q=0.0, alpha_inf=100.0, t_star=20.0, sigma_c=50.0, m=10.0)  # test_crystal_visco_plastic.jl:46
CRYSTAL_RODRIGUES = (-0.665267, 0.0203875, 1.08633)                                                      # test_crystal_visco_plastic.jl:45
NEOHOOK_PARAMS = dict(lambda_=1.0, mu=1.0, c2=0.0, c3=0.0)                                               # test_neohook.jl:2
YEOH_PARAMS = dict(lambda_=1.0, mu=1.0, c2=-1.0, c3=1.0)                                                 # test_yeoh.jl:3
XU_PARAMS = dict(sigma_max=400.0, tau_max=200.0, delta_n=0.002, delta_t=0.002, Delta_n_star=0.01)        # test_xu_needleman.jl:2-12

# ---- input ranges per config: (nc, lo, hi, seed) -----------------------------------------------
PLASTIC_AMPL = 6.6e-4      # ~50 % of points yield from the virgin state
CRYSTAL_AMPL = 3.0e-3


def c1_strain():
    return 6, -2e-3, 2e-3, SEED0 + 1


def c2_strain(step):
    """step 0 = A (from the zero state), step 1 = B (fed with A's state; the timed step)."""
    return 6, -PLASTIC_AMPL, PLASTIC_AMPL, SEED0 + 2 + 1000 * step


def c3_defgrad():
    lo = np.full(9, -0.2)
    hi = np.full(9, 0.2)
    for d in (0, 4, 8):
        lo[d] += 1.0
        hi[d] += 1.0
    return 9, lo, hi, SEED0 + 3


def c4_strain_increment(step):
    return 6, -CRYSTAL_AMPL, CRYSTAL_AMPL, SEED0 + 4 + 1000 * step


def c5_jump():
    lo = np.array([-0.004, -0.004, -0.001])
    hi = np.array([0.004, 0.004, 0.008])
    return 3, lo, hi, SEED0 + 5
