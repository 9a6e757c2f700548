// Device-pointer entry points of the streaming (HBM-bound) models: LinearElastic, Plastic (J2),
// NeoHook / Yeoh / StVenant, XuNeedleman, plus the synthetic-input generator.
#include <cuda_runtime.h>
#include <string.h>
#include "../../include/mmb200.h"
#include "mm_kernels.cuh"

namespace mm {

int* ops_tuning_slot(const char* name);      // run-time switches of this translation unit (below)

// ---- Ops ----------------------------------------------------------------------------------------
struct ElasticOp {
    static constexpr int NIN = 1, NOUT = 2, SCRATCH = 0;
    __host__ __device__ static constexpr int in_nc(int) { return 6; }
    __host__ __device__ static constexpr int out_nc(int a) { return a == 0 ? 6 : 36; }
    ElasticK k;
    template <class Acc> __device__ __forceinline__ void point(Acc& acc, int64_t, const Scratch&) const { linear_elastic_point(k, acc); }
};

template <> struct SoaTangentCap<ElasticOp> { static constexpr int blocks = 2; };
// the tangent-writing launches gain from the hint (profiles/r02_stream_prefetch_ab.jsonl, 1e8 points: LinearElastic 6.03 -> 5.91 ms, NeoHook / Yeoh
// 6.48-6.65 -> 6.29-6.36 ms, XuNeedleman 1.83-1.95 -> 1.82-1.84 ms); the read-dominated σ-only launch loses (1.42 -> 1.46 ms) and gets none
inline int stream_prefetch_dist(bool writes_tangent, int dflt) {
    const int forced = *ops_tuning_slot("stream_soa_prefetch");
    return forced >= 0 ? forced : (writes_tangent ? dflt : 0);
}
template <> struct SoaPrefetch<ElasticOp> { static int dist(const Ptrs<1, 2>& p) { return stream_prefetch_dist(p.out[1] != nullptr, 65536); } };

struct PlasticOp {
    static constexpr int NIN = 2, NOUT = 3, SCRATCH = 0;
    __host__ __device__ static constexpr int in_nc(int a) { return a == 0 ? 6 : 14; }
    __host__ __device__ static constexpr int out_nc(int a) { return a == 0 ? 6 : (a == 1 ? 36 : 14); }
    PlasticK k;
    uint8_t* flags;
    uint16_t* iters;
    int32_t* n_fail;
    // parts_mode (the producer half of the compact return, mm_compact.cu): output 1 receives the tangent in its FACTORED form
    // {gξ, u(6), n(6)} — components 0..12 of the 36 — instead of the 36 entries.  A run-time switch inside the ONE kernel rather than
    // a second instantiation on purpose: the σ / state arithmetic (and nvcc's fma contraction choices in it) are then the very same
    // machine code for the dense and the compact return, which is what makes the two bit-identical.
    int parts_mode;
    template <class Acc> __device__ __forceinline__ void point(Acc& acc, int64_t i, const Scratch&) const {
        int it = 0;
        PlasticTangentParts tp;
        const int flag = plastic_point_parts(k, acc, &it, tp);
        if (acc.template has_out<1>()) {
            if (parts_mode) {
                acc.template out<1>(0, tp.gxi);
#pragma unroll
                for (int c = 0; c < 6; ++c) { acc.template out<1>(1 + c, tp.u[c]); acc.template out<1>(7 + c, tp.n[c]); }
            } else {
#pragma unroll
                for (int J = 0; J < 6; ++J)
#pragma unroll
                    for (int I = 0; I < 6; ++I) acc.template out<1>(I + 6 * J, plastic_tangent_entry(k, tp, I, J));
            }
        }
        if (flags) flags[i] = (uint8_t)flag;
        if (iters) iters[i] = (uint16_t)(it > 65535 ? 65535 : it);
        if ((flag & MM_FLAG_FAILED) && n_fail) atomicAdd(n_fail, 1);
    }
};
template <> struct SoaPrefetch<PlasticOp> { static int dist(const Ptrs<2, 3>&) { return *ops_tuning_slot("plastic_soa_prefetch"); } };
template <> struct OutAlias<PlasticOp> { __host__ __device__ static constexpr int of(int a) { return a == 0 ? 0 : (a == 2 ? 1 : -1); } };   // σ <- ε slots, state <- state slots

#ifndef MM_PLASTIC_SOA_BS
#define MM_PLASTIC_SOA_BS 128
#endif
#ifndef MM_PLASTIC_AOS_BS
#define MM_PLASTIC_AOS_BS 64
#endif
#ifndef MM_STREAM_AOS_BS
#define MM_STREAM_AOS_BS 128
#endif
#ifndef MM_SMALL_AOS_BS
#define MM_SMALL_AOS_BS 512       // AoS tile of the ops with <= 20 doubles per point (XuNeedleman, energies): 128-point tiles are 3-9 KB copies, sync-bound (Xu 2.08 -> 1.93 ms, energies 1.30 -> 1.15 ms)
#endif
#ifndef MM_PLASTIC_MINBLOCKS
#define MM_PLASTIC_MINBLOCKS 3
#endif
#ifndef MM_PLASTIC_MINBLOCK_THREADS
#define MM_PLASTIC_MINBLOCK_THREADS 128
#endif
#ifndef MM_PLASTIC_AOS_MINBLOCKS
#define MM_PLASTIC_AOS_MINBLOCKS 2
#endif
template <> struct MinBlocksAos<PlasticOp> { static constexpr int value = MM_PLASTIC_AOS_MINBLOCKS, block = 128; };   // 256 threads -> 238 registers, no spills: 10.1 ms vs 11.2 ms at the SoA kernel's 168
template <> struct MinBlocks<PlasticOp> { static constexpr int value = MM_PLASTIC_MINBLOCKS, block = MM_PLASTIC_MINBLOCK_THREADS; };   // 128 threads x 3 blocks/SM -> <= 168 regs: measured best of {1,2,3,4,5} (profiles/r01_plastic_regcap_sweep.log, r01_plastic_aos_regcap_sweep.log)

template <int MODEL, int KIND, int TAN> struct HyperOp {
    static constexpr int NIN = 1, NOUT = 2, SCRATCH = 0;
    __host__ __device__ static constexpr int in_nc(int) { return KIND == 1 ? 9 : 6; }
    __host__ __device__ static constexpr int out_nc(int a) { return a == 0 ? (TAN == 2 ? 9 : 6) : (TAN == 2 ? 81 : 36); }
    HyperK k;
    template <class Acc> __device__ __forceinline__ void point(Acc& acc, int64_t, const Scratch&) const { hyper_point_ex<MODEL, KIND, TAN>(k, acc); }
};

template <int MODEL, int KIND, int TAN> struct SoaPrefetch<HyperOp<MODEL, KIND, TAN>> { static int dist(const Ptrs<1, 2>& p) { return stream_prefetch_dist(p.out[1] != nullptr, 32768); } };
template <int MODEL, int KIND, int TAN> struct SoaTangentCap<HyperOp<MODEL, KIND, TAN>> { static constexpr int blocks = TAN == 1 ? 0 : 2; };   // ∂S∂E measured slower capped (6.82 vs 6.66 ms)

template <int DIM> struct XuOp {
    static constexpr int NIN = 1, NOUT = 2, SCRATCH = 0;
    __host__ __device__ static constexpr int in_nc(int) { return DIM; }
    __host__ __device__ static constexpr int out_nc(int a) { return a == 0 ? DIM : DIM * DIM; }
    XuK k;
    template <class Acc> __device__ __forceinline__ void point(Acc& acc, int64_t, const Scratch&) const { xu_point<DIM>(k, acc); }
};

#ifndef MM_XU_MINBLOCKS
#define MM_XU_MINBLOCKS 1          // occupancy target of the SoA kernel in blocks of MM_XU_SOA_BS threads (1 = no register cap)
#endif
#ifndef MM_XU_SOA_BS
#define MM_XU_SOA_BS 256
#endif
template <int DIM> struct SoaPrefetch<XuOp<DIM>> { static int dist(const Ptrs<1, 2>& p) { return stream_prefetch_dist(p.out[1] != nullptr, 65536); } };
template <int DIM> struct MinBlocks<XuOp<DIM>> { static constexpr int value = MM_XU_MINBLOCKS, block = MM_XU_SOA_BS; };
template <int DIM> struct MinBlocksAos<XuOp<DIM>> { static constexpr int value = 1, block = 32; };
#ifndef MM_XU_SOA_TILED
#define MM_XU_SOA_TILED 1
#endif
#ifndef MM_XU_TILE_BS
#define MM_XU_TILE_BS 256
#endif
template <int DIM> struct SoaTiled<XuOp<DIM>> { static constexpr bool value = MM_XU_SOA_TILED != 0; static constexpr int BS = MM_XU_TILE_BS; };
#ifndef MM_XU_CAP
#define MM_XU_CAP 0
#endif
template <int DIM> struct SoaTangentCap<XuOp<DIM>> { static constexpr int blocks = MM_XU_CAP; };

struct TransvIsoOp {
    static constexpr int NIN = 2, NOUT = 2, SCRATCH = 0;
    __host__ __device__ static constexpr int in_nc(int a) { return a == 0 ? 6 : 3; }
    __host__ __device__ static constexpr int out_nc(int a) { return a == 0 ? 6 : 36; }
    TransvIsoK k;
    template <class Acc> __device__ __forceinline__ void point(Acc& acc, int64_t, const Scratch&) const { transviso_point(k, acc); }
};
struct ElasticEnergyOp {
    static constexpr int NIN = 1, NOUT = 1, SCRATCH = 0;
    __host__ __device__ static constexpr int in_nc(int) { return 6; }
    __host__ __device__ static constexpr int out_nc(int) { return 1; }
    ElasticK k;
    template <class Acc> __device__ __forceinline__ void point(Acc& acc, int64_t, const Scratch&) const { linear_elastic_energy_point(k, acc); }
};
template <int MODEL, int KIND> struct HyperEnergyOp {
    static constexpr int NIN = 1, NOUT = 1, SCRATCH = 0;
    __host__ __device__ static constexpr int in_nc(int) { return KIND == 1 ? 9 : 6; }
    __host__ __device__ static constexpr int out_nc(int) { return 1; }
    HyperK k;
    template <class Acc> __device__ __forceinline__ void point(Acc& acc, int64_t, const Scratch&) const { hyper_energy_point<MODEL, KIND>(k, acc); }
};

// ---- synthetic inputs ---------------------------------------------------------------------------
struct SynthParams { double lo[48], hi[48]; };
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void __launch_bounds__(256) synth_kernel(double* out, int64_t N, int nc, int layout, uint64_t seed, const SynthParams sp) {
    const int64_t total = N * nc;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
        int64_t i; int c;
        if (layout == MM_LAYOUT_SOA) { c = (int)(e / N); i = e - (int64_t)c * N; }
        else                         { i = e / nc; c = (int)(e - i * nc); }
        const uint64_t h = splitmix64(seed + (uint64_t)i * (uint64_t)nc + (uint64_t)c);
        const double u = __ull2double_rn(h >> 11) * (1.0 / 9007199254740992.0);
        // no FMA contraction: bit-identical to the host generator used by the parity harness
        out[e] = __dadd_rn(sp.lo[c], __dmul_rn(__dsub_rn(sp.hi[c], sp.lo[c]), u));
    }
}

// the 81-entry ∂Pᵀ∂F route needs ~250 registers unconstrained (one 256-thread block per SM, 8 warps: 0.84 of the HBM peak);
// capped at 168 (3 x 128 threads per SM) it streams like the others
#ifndef MM_DPDF_MINBLOCKS
#define MM_DPDF_MINBLOCKS 3
#endif
#ifndef MM_DPDF_AOS_MINBLOCKS
#define MM_DPDF_AOS_MINBLOCKS MM_DPDF_MINBLOCKS
#endif
template <int MODEL, int KIND> struct MinBlocksAos<HyperOp<MODEL, KIND, 2>> { static constexpr int value = MM_DPDF_AOS_MINBLOCKS, block = 128; };
template <int MODEL, int KIND> struct MinBlocks<HyperOp<MODEL, KIND, 2>> { static constexpr int value = MM_DPDF_MINBLOCKS, block = 128; };

template <int MODEL, int KIND, int TAN>
int launch_hyper_mkt(const HyperK& k, const double* strain, double* S, double* T, int64_t N, int layout, cudaStream_t st) {
    Ptrs<1, 2> p;
    p.in[0] = strain; p.out[0] = S; p.out[1] = T;
    HyperOp<MODEL, KIND, TAN> op; op.k = k;
    return launch_points<HyperOp<MODEL, KIND, TAN>, (TAN == 2 ? 128 : 256), (TAN == 2 ? 64 : MM_STREAM_AOS_BS)>(op, p, N, layout, st, "mm_hyper");
}
template <int MODEL> int launch_hyper(const HyperK& k, const double* strain, int strain_kind, int tangent_kind, double* S, double* T, int64_t N, int layout, cudaStream_t st) {
    if (tangent_kind == MM_TANGENT_DPDF) return launch_hyper_mkt<MODEL, 1, 2>(k, strain, S, T, N, layout, st);
    if (tangent_kind == MM_TANGENT_DSDE) {
        if (strain_kind == MM_STRAIN_F) return launch_hyper_mkt<MODEL, 1, 1>(k, strain, S, T, N, layout, st);
        if (strain_kind == MM_STRAIN_E) return launch_hyper_mkt<MODEL, 2, 1>(k, strain, S, T, N, layout, st);
        return launch_hyper_mkt<MODEL, 0, 1>(k, strain, S, T, N, layout, st);
    }
    if (strain_kind == MM_STRAIN_F) return launch_hyper_mkt<MODEL, 1, 0>(k, strain, S, T, N, layout, st);
    if (strain_kind == MM_STRAIN_E) return launch_hyper_mkt<MODEL, 2, 0>(k, strain, S, T, N, layout, st);
    return launch_hyper_mkt<MODEL, 0, 0>(k, strain, S, T, N, layout, st);
}

template <int MODEL> int launch_hyper_energy(const HyperK& k, const double* strain, int strain_kind, double* psi, int64_t N, int layout, cudaStream_t st) {
    Ptrs<1, 1> p; p.in[0] = strain; p.out[0] = psi;
    if (strain_kind == MM_STRAIN_F) { HyperEnergyOp<MODEL, 1> op; op.k = k; return launch_points<HyperEnergyOp<MODEL, 1>, 256, MM_SMALL_AOS_BS>(op, p, N, layout, st, "mm_hyper_energy"); }
    if (strain_kind == MM_STRAIN_E) { HyperEnergyOp<MODEL, 2> op; op.k = k; return launch_points<HyperEnergyOp<MODEL, 2>, 256, MM_SMALL_AOS_BS>(op, p, N, layout, st, "mm_hyper_energy"); }
    HyperEnergyOp<MODEL, 0> op; op.k = k;
    return launch_points<HyperEnergyOp<MODEL, 0>, 256, MM_SMALL_AOS_BS>(op, p, N, layout, st, "mm_hyper_energy");
}

// internal (mm_compact.cu): the J2 update with output 1 in factored form — SoA: parts[c*N + i], c < 13; AoS: parts[i*36 + c] (the tile
// kernel stages output 1 with its compile-time stride of 36; only the first 13 slots of a point are meaningful)
int launch_plastic_parts(const MMPlastic* p, const MMNewtonOpts* opts, const double* eps, const double* state_in, double* sig, double* parts, double* state_out,
                         uint8_t* flags, uint16_t* iters, int32_t* n_fail, int64_t N, int layout, cudaStream_t st) {
    PlasticOp op;
    op.k = make_plastic_k(p->E, p->nu, p->sigma_y, p->H, p->r, p->kappa_inf, p->alpha_inf, opts ? opts->tol : 1e-8, opts ? opts->max_iter : 1000);
    op.flags = flags; op.iters = iters; op.n_fail = n_fail; op.parts_mode = 1;
    Ptrs<2, 3> ptrs; ptrs.in[0] = eps; ptrs.in[1] = state_in; ptrs.out[0] = sig; ptrs.out[1] = parts; ptrs.out[2] = state_out;
    return launch_points<PlasticOp, MM_PLASTIC_SOA_BS, MM_PLASTIC_AOS_BS>(op, ptrs, N, layout, st, "mm_plastic_compact");
}

// run-time A/B switches of this translation unit (mm_set_tuning): "xu_soa_tiled" 0 (default) = direct SoA kernel, 1 = copy-engine tiles;
// "plastic_soa_prefetch" = how many points ahead the J2 SoA kernel asks L2 for its input lines (0 = no hint, default 28 416);
// "stream_soa_prefetch" = the same for the LinearElastic / hyperelastic / XuNeedleman SoA kernels (-1 = per-op defaults for the tangent-writing launches)
int* ops_tuning_slot(const char* name) {
    static int xu_soa_tiled = [] { const char* e = getenv("MM_XU_SOA_TILED_RT"); return (e && e[0] == '1') ? 1 : 0; }();
    if (!strcmp(name, "xu_soa_tiled")) return &xu_soa_tiled;
    // measured on 1e8 points, alternating rounds on two boxes (profiles/r02_plastic_prefetch_ab*.jsonl): back to back 10.33-10.75 ms without the
    // hint, 9.90-10.29 ms with it anywhere from 14 208 to 56 832 points ahead (1/4 to 1 resident wave of 148 x 3 blocks), 10.2 ms at 113 664,
    // 10.9 ms at 227 328 (the lines are evicted again before they are used), same bits
    static int plastic_soa_prefetch = [] { const char* e = getenv("MM_PLASTIC_SOA_PREFETCH"); return e ? atoi(e) : 28416; }();
    if (!strcmp(name, "plastic_soa_prefetch")) return &plastic_soa_prefetch;
    static int stream_soa_prefetch = [] { const char* e = getenv("MM_STREAM_SOA_PREFETCH"); return e ? atoi(e) : -1; }();     // LinearElastic, hyperelastic, XuNeedleman SoA kernels: -1 = per-op defaults, >= 0 = that distance for all of them (0 = no hint)
    if (!strcmp(name, "stream_soa_prefetch")) return &stream_soa_prefetch;
    return nullptr;
}

}  // namespace mm

using namespace mm;

extern "C" {

int mm_linear_elastic(const MMLinearElastic* p, const double* eps, double* sig, double* tan_or_null, int64_t N, int layout, void* stream) {
    if (!p || N < 0 || (N > 0 && (!eps || !sig))) return set_error(MM_ERR_ARG, "mm_linear_elastic: bad argument");
    ElasticOp op; op.k = make_elastic_k(p->E, p->nu);
    Ptrs<1, 2> ptrs; ptrs.in[0] = eps; ptrs.out[0] = sig; ptrs.out[1] = tan_or_null;
    return launch_points<ElasticOp, 256, MM_STREAM_AOS_BS>(op, ptrs, N, layout, (cudaStream_t)stream, "mm_linear_elastic");
}

int mm_plastic(const MMPlastic* p, const double* eps, const double* state_in, double* sig, double* tan, double* state_out,
               uint8_t* flags, uint16_t* iters, int32_t* n_fail, const MMNewtonOpts* opts, int64_t N, int layout, void* stream) {
    if (!p || N < 0 || (N > 0 && (!eps || !state_in || !sig || !state_out))) return set_error(MM_ERR_ARG, "mm_plastic: bad argument");
    PlasticOp op;
    op.k = make_plastic_k(p->E, p->nu, p->sigma_y, p->H, p->r, p->kappa_inf, p->alpha_inf, opts ? opts->tol : 1e-8, opts ? opts->max_iter : 1000);
    op.flags = flags; op.iters = iters; op.n_fail = n_fail; op.parts_mode = 0;
    Ptrs<2, 3> ptrs; ptrs.in[0] = eps; ptrs.in[1] = state_in; ptrs.out[0] = sig; ptrs.out[1] = tan; ptrs.out[2] = state_out;
    return launch_points<PlasticOp, MM_PLASTIC_SOA_BS, MM_PLASTIC_AOS_BS>(op, ptrs, N, layout, (cudaStream_t)stream, "mm_plastic");
}

int mm_hyper_ex(int model, const MMHyper* p, const double* strain, int strain_kind, int tangent_kind, double* stress, double* tangent,
                int64_t N, int layout, void* stream) {
    if (!p || N < 0 || (N > 0 && (!strain || !stress))) return set_error(MM_ERR_ARG, "mm_hyper: bad argument");
    if (strain_kind != MM_STRAIN_C && strain_kind != MM_STRAIN_F && strain_kind != MM_STRAIN_E) return set_error(MM_ERR_ARG, "mm_hyper: unknown strain_kind %d", strain_kind);
    if (tangent_kind != MM_TANGENT_DSDC && tangent_kind != MM_TANGENT_DSDE && tangent_kind != MM_TANGENT_DPDF) return set_error(MM_ERR_ARG, "mm_hyper: unknown tangent_kind %d", tangent_kind);
    if (tangent_kind == MM_TANGENT_DPDF && strain_kind != MM_STRAIN_F)
        return set_error(MM_ERR_ARG, "mm_hyper: the dPdF tangent needs the deformation gradient as input (src/traits.jl:67-72)");
    HyperK k; k.lambda = p->lambda; k.mu = p->mu; k.c2 = p->c2; k.c3 = p->c3;
    cudaStream_t st = (cudaStream_t)stream;
    switch (model) {
        case MM_NEOHOOK: return launch_hyper<0>(k, strain, strain_kind, tangent_kind, stress, tangent, N, layout, st);
        case MM_YEOH: return launch_hyper<1>(k, strain, strain_kind, tangent_kind, stress, tangent, N, layout, st);
        case MM_STVENANT: return launch_hyper<2>(k, strain, strain_kind, tangent_kind, stress, tangent, N, layout, st);
    }
    return set_error(MM_ERR_ARG, "mm_hyper: unknown model %d", model);
}

int mm_hyper(int model, const MMHyper* p, const double* strain, int strain_kind, double* S, double* dSdC, int64_t N, int layout, void* stream) {
    return mm_hyper_ex(model, p, strain, strain_kind, MM_TANGENT_DSDC, S, dSdC, N, layout, stream);
}

int mm_xu_needleman(const MMXuNeedleman* p, const double* jump, double* T, double* dTdD, int dim, int64_t N, int layout, void* stream) {
    if (!p || N < 0 || (N > 0 && (!jump || !T))) return set_error(MM_ERR_ARG, "mm_xu_needleman: bad argument");
    const XuK k = make_xu_k(p->Phi_n, p->Phi_t, p->delta_n, p->delta_t, p->Delta_n_star);
    Ptrs<1, 2> ptrs; ptrs.in[0] = jump; ptrs.out[0] = T; ptrs.out[1] = dTdD;
    const bool tiled = *ops_tuning_slot("xu_soa_tiled") != 0;
    if (dim == 3) { XuOp<3> op; op.k = k; return launch_points<XuOp<3>, MM_XU_SOA_BS, MM_SMALL_AOS_BS>(op, ptrs, N, layout, (cudaStream_t)stream, "mm_xu_needleman", tiled); }
    if (dim == 2) { XuOp<2> op; op.k = k; return launch_points<XuOp<2>, MM_XU_SOA_BS, MM_SMALL_AOS_BS>(op, ptrs, N, layout, (cudaStream_t)stream, "mm_xu_needleman", tiled); }
    return set_error(MM_ERR_ARG, "mm_xu_needleman: dim must be 2 or 3 (got %d)", dim);
}

int mm_transversely_isotropic_setup(double E_L, double E_T, double G_LT, double nu_LT, double nu_TT, MMTransverselyIsotropic* out) {
    if (!out) return set_error(MM_ERR_ARG, "mm_transversely_isotropic_setup: null output");
    const double den = 2 * E_T * nu_LT * nu_LT - E_L + E_L * nu_TT;                     // transverselyisotropic.jl:46-50
    out->M_par = (E_L * E_L * (nu_TT - 1)) / den;
    out->L_par = -(E_L * E_T * nu_LT) / den;
    out->L_perp = -(E_T * (E_T * nu_LT * nu_LT + E_L * nu_TT)) / ((nu_TT + 1) * den);
    out->G_perp = E_T / (2 * (nu_TT + 1));
    out->G_par = G_LT;
    return MM_OK;
}

int mm_transversely_isotropic(const MMTransverselyIsotropic* p, const double* eps, const double* a3, double* sig, double* tan, int64_t N, int layout, void* stream) {
    if (!p || N < 0 || (N > 0 && (!eps || !a3 || !sig))) return set_error(MM_ERR_ARG, "mm_transversely_isotropic: bad argument");
    TransvIsoOp op; op.k.L_perp = p->L_perp; op.k.L_par = p->L_par; op.k.M_par = p->M_par; op.k.G_perp = p->G_perp; op.k.G_par = p->G_par;
    Ptrs<2, 2> ptrs; ptrs.in[0] = eps; ptrs.in[1] = a3; ptrs.out[0] = sig; ptrs.out[1] = tan;
    return launch_points<TransvIsoOp, 256, MM_STREAM_AOS_BS>(op, ptrs, N, layout, (cudaStream_t)stream, "mm_transversely_isotropic");
}

int mm_linear_elastic_energy(const MMLinearElastic* p, const double* eps, double* psi, int64_t N, int layout, void* stream) {
    if (!p || N < 0 || (N > 0 && (!eps || !psi))) return set_error(MM_ERR_ARG, "mm_linear_elastic_energy: bad argument");
    ElasticEnergyOp op; op.k = make_elastic_k(p->E, p->nu);
    Ptrs<1, 1> ptrs; ptrs.in[0] = eps; ptrs.out[0] = psi;
    return launch_points<ElasticEnergyOp, 256, MM_SMALL_AOS_BS>(op, ptrs, N, layout, (cudaStream_t)stream, "mm_linear_elastic_energy");
}

int mm_hyper_energy(int model, const MMHyper* p, const double* strain, int strain_kind, double* psi, int64_t N, int layout, void* stream) {
    if (!p || N < 0 || (N > 0 && (!strain || !psi))) return set_error(MM_ERR_ARG, "mm_hyper_energy: bad argument");
    if (strain_kind != MM_STRAIN_C && strain_kind != MM_STRAIN_F && strain_kind != MM_STRAIN_E) return set_error(MM_ERR_ARG, "mm_hyper_energy: unknown strain_kind %d", strain_kind);
    HyperK k; k.lambda = p->lambda; k.mu = p->mu; k.c2 = p->c2; k.c3 = p->c3;
    cudaStream_t st = (cudaStream_t)stream;
    switch (model) {
        case MM_NEOHOOK: return launch_hyper_energy<0>(k, strain, strain_kind, psi, N, layout, st);
        case MM_YEOH: return launch_hyper_energy<1>(k, strain, strain_kind, psi, N, layout, st);
        case MM_STVENANT: return launch_hyper_energy<2>(k, strain, strain_kind, psi, N, layout, st);
    }
    return set_error(MM_ERR_ARG, "mm_hyper_energy: unknown model %d", model);
}

int mm_synth_uniform(double* out, int64_t N, int nc, int layout, uint64_t seed, const double* lo, const double* hi, void* stream) {
    if (!out || !lo || !hi || N < 0 || nc < 1 || nc > 48) return set_error(MM_ERR_ARG, "mm_synth_uniform: bad argument");
    if (N == 0) return MM_OK;
    SynthParams sp;
    for (int c = 0; c < nc; ++c) { sp.lo[c] = lo[c]; sp.hi[c] = hi[c]; }
    int64_t blocks = (N * nc + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    synth_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(out, N, nc, layout, seed, sp);
    return check_launch("mm_synth_uniform");
}

}  // extern "C"
