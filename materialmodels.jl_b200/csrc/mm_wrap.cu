// Device-pointer entry points of the dimension wrappers (reference src/wrappers.jl) around LinearElastic and Plastic.
#include <cuda_runtime.h>
#include "../../include/mmb200.h"
#include "mm_kernels.cuh"
#include "mm_crystal_math.cuh"
#include "mm_wrap_math.cuh"
#include "mm_wrap_lp.cuh"
#include <string.h>

namespace mm {

template <int KIND> struct WrappedElasticOp {
    static constexpr int NIN = 1, NOUT = 2, SCRATCH = 0;
    __host__ __device__ static constexpr int in_nc(int) { return wrap_nc(KIND); }
    __host__ __device__ static constexpr int out_nc(int a) { return a == 0 ? wrap_nc(KIND) : wrap_nc(KIND) * wrap_nc(KIND); }
    ElasticK k;
    double tol; int max_iter;
    uint8_t* outer_iters;
    int32_t* n_fail;
    template <class Acc> __device__ __forceinline__ void point(Acc& acc, int64_t i, const Scratch&) const {
        int oit = 0, iit = 0;
        const ElasticK kk = k;
        const int flag = wrapped_point<KIND, 0>([&](RegIO<0>& io, int*) { linear_elastic_point(kk, io); return 0; }, acc, tol, max_iter, &oit, &iit);
        if (outer_iters) outer_iters[i] = (uint8_t)oit;
        if (flag && n_fail) atomicAdd(n_fail, 1);
    }
};

struct LazyJ2IO {
    double eps[6], st_in[14], sig[6], st_out[14];
    PlasticK k;
    PlasticTangentParts tp;
    template <int A> __device__ __forceinline__ double in(int c) const { return A == 0 ? eps[c] : st_in[c]; }
    template <int A> __device__ __forceinline__ void out(int c, double v) { if (A == 0) sig[c] = v; else if (A == 2) st_out[c] = v; }
    template <int A> __device__ __forceinline__ bool has_out() const { return A != 1; }
    __device__ __forceinline__ double T(int idx) const { return plastic_tangent_entry(k, tp, idx % 6, idx / 6); }
};
template <int KIND> struct WrappedPlasticOp {
    static constexpr int NIN = 2, NOUT = 3, SCRATCH = 0;
    __host__ __device__ static constexpr int in_nc(int a) { return a == 0 ? wrap_nc(KIND) : 14; }
    __host__ __device__ static constexpr int out_nc(int a) { return a == 0 ? wrap_nc(KIND) : (a == 1 ? wrap_nc(KIND) * wrap_nc(KIND) : 14); }
    PlasticK k;
    double tol; int max_iter;
    uint8_t* flags;
    uint16_t* iters;
    uint8_t* outer_iters;
    int32_t* n_fail;
    template <class Acc> __device__ __forceinline__ void point(Acc& acc, int64_t i, const Scratch&) const {
        int oit = 0, iit = 0;
        // the inner model leaves the three ingredients of its tangent (13 doubles) instead of 36 entries; the wrapper evaluates the 1-25
        // entries of the zero-stress block per outer pass and the rest once, at the converged pass (same expression, same bits)
        LazyJ2IO io;
        io.k = k;
#pragma unroll
        for (int c = 0; c < 14; ++c) io.st_in[c] = acc.template in<1>(c);
        const int flag = wrapped_core<KIND>([&](LazyJ2IO& o, int* it) { return plastic_point_parts(o.k, o, it, o.tp); }, acc, io, tol, max_iter, &oit, &iit);
#pragma unroll
        for (int c = 0; c < 14; ++c) acc.template out<2>(c, io.st_out[c]);
        if (flags) flags[i] = (uint8_t)flag;
        if (iters) iters[i] = (uint16_t)(iit > 65535 ? 65535 : iit);
        if (outer_iters) outer_iters[i] = (uint8_t)oit;
        if ((flag & 0xC0) && n_fail) atomicAdd(n_fail, 1);
    }
};
template <int KIND> struct OutAlias<WrappedPlasticOp<KIND>> { __host__ __device__ static constexpr int of(int a) { return a == 0 ? 0 : (a == 2 ? 1 : -1); } };
#ifndef MM_WRAP_MINBLOCKS
#define MM_WRAP_MINBLOCKS 3
#endif
#ifndef MM_WRAP_AOS_MINBLOCKS
#define MM_WRAP_AOS_MINBLOCKS MM_WRAP_MINBLOCKS
#endif
template <int KIND> struct MinBlocksAos<WrappedPlasticOp<KIND>> { static constexpr int value = MM_WRAP_AOS_MINBLOCKS, block = 128; };
template <int KIND> struct MinBlocks<WrappedPlasticOp<KIND>> { static constexpr int value = MM_WRAP_MINBLOCKS, block = 128; };

// wrappers.jl is generic in the material; TransverselyIsotropic (the laminate-ply model: ShellStress / PlaneStress use) carries
// its direction a3 as the (unchanged) state, so the inner model reads it through in<1> and nothing is written to out<2>
template <class Acc> struct NoStateOut {
    Acc& a;
    template <int A> __device__ __forceinline__ double in(int c) const { return a.template in<A>(c); }
    template <int A> __device__ __forceinline__ void out(int c, double v) { if (A < 2) a.template out<(A < 2 ? A : 0)>(c, v); }
    template <int A> __device__ __forceinline__ bool has_out() const { return A < 2 ? a.template has_out<(A < 2 ? A : 0)>() : false; }
};
template <int KIND> struct WrappedTransvIsoOp {
    static constexpr int NIN = 2, NOUT = 2, SCRATCH = 0;
    __host__ __device__ static constexpr int in_nc(int a) { return a == 0 ? wrap_nc(KIND) : 3; }
    __host__ __device__ static constexpr int out_nc(int a) { return a == 0 ? wrap_nc(KIND) : wrap_nc(KIND) * wrap_nc(KIND); }
    TransvIsoK k;
    double tol; int max_iter;
    uint8_t* outer_iters;
    int32_t* n_fail;
    template <class Acc> __device__ __forceinline__ void point(Acc& acc, int64_t i, const Scratch&) const {
        int oit = 0, iit = 0;
        const TransvIsoK kk = k;
        NoStateOut<Acc> na{acc};
        const int flag = wrapped_point<KIND, 3>([&](RegIO<3>& io, int*) { transviso_point(kk, io); return 0; }, na, tol, max_iter, &oit, &iit);
        if (outer_iters) outer_iters[i] = (uint8_t)oit;
        if (flag && n_fail) atomicAdd(n_fail, 1);
    }
};
#ifndef MM_WRAPTI_MINBLOCKS
#define MM_WRAPTI_MINBLOCKS 4
#endif
template <int KIND> struct MinBlocks<WrappedTransvIsoOp<KIND>> { static constexpr int value = MM_WRAPTI_MINBLOCKS, block = 128; };
template <int KIND>
int launch_wrapped_transviso(const MMTransverselyIsotropic* p, const double* eps, const double* a3, double* sig, double* tan, uint8_t* oit, int32_t* n_fail,
                             const MMNewtonOpts* w, int64_t N, int layout, cudaStream_t st) {
    WrappedTransvIsoOp<KIND> op;
    op.k.L_perp = p->L_perp; op.k.L_par = p->L_par; op.k.M_par = p->M_par; op.k.G_perp = p->G_perp; op.k.G_par = p->G_par;
    op.tol = w ? w->tol : 1e-8; op.max_iter = w ? w->max_iter : 10; op.outer_iters = oit; op.n_fail = n_fail;
    Ptrs<2, 2> ptrs; ptrs.in[0] = eps; ptrs.in[1] = a3; ptrs.out[0] = sig; ptrs.out[1] = tan;
    return launch_points<WrappedTransvIsoOp<KIND>, 128, 128>(op, ptrs, N, layout, st, "mm_wrapped_transversely_isotropic");
}

// CrystalViscoPlastic under a dimension wrapper (wrappers.jl is generic in the material): the inner model is the structured Newton
// (crystal_point_fast, per-system vectors in the thread's shared-memory scratch column) on the register accessor, with the general
// pivoted 12x12 path as its fallback.  Functional, not tuned: every outer update is a full inner crystal solve including its tangent.  `active` bit 13 = outer Newton failed, bit 15 = inner solve failed.
#define MM_ACTIVE_WRAPPER_FAILED 0x2000u
template <int KIND> struct WrappedCrystalOp {
    static constexpr int NIN = 2, NOUT = 3, SCRATCH = MM_CRYSTAL_SCRATCH;
    __host__ __device__ static constexpr int in_nc(int a) { return a == 0 ? wrap_nc(KIND) : 42; }
    __host__ __device__ static constexpr int out_nc(int a) { return a == 0 ? wrap_nc(KIND) : (a == 1 ? wrap_nc(KIND) * wrap_nc(KIND) : 42); }
    CrystalK k;
    double tol; int max_iter;
    uint16_t* active;
    uint16_t* iters;
    uint8_t* outer_iters;
    int32_t* n_fail;
    template <class Acc> __device__ __forceinline__ void point(Acc& acc, int64_t i, const Scratch& scr) const {
        int oit = 0, iit = 0;
        unsigned mask = 0;
        const int flag = wrapped_point<KIND, 42>(
            [&](RegIO<42>& io, int* it) {
                bool fb = !k.h_struct;
                if (!fb) mask = (k.b_h != 0.0) ? crystal_point_fast<true>(k, io, scr, it, &fb) : crystal_point_fast<false>(k, io, scr, it, &fb);
                if (fb) mask = crystal_point(k, io, it);               // general pivoted path, Jacobian in local memory (rare)
                return (mask & MM_ACTIVE_FAILED) ? 0x80 : 0;
            },
            acc, tol, max_iter, &oit, &iit);
        if (flag & 0x40) mask |= MM_ACTIVE_WRAPPER_FAILED;
        if (active) active[i] = (uint16_t)mask;
        if (iters) iters[i] = (uint16_t)(iit > 65535 ? 65535 : iit);
        if (outer_iters) outer_iters[i] = (uint8_t)oit;
        if ((flag & 0xC0) && n_fail) atomicAdd(n_fail, 1);
    }
};
template <int KIND>
int launch_wrapped_crystal(const CrystalK& k, const double* deps, const double* sin, double* sig, double* tan, double* sout, uint16_t* active, uint16_t* iters,
                           uint8_t* oit, int32_t* n_fail, const MMNewtonOpts* w, int64_t N, int layout, cudaStream_t st) {
    WrappedCrystalOp<KIND> op;
    op.k = k; op.tol = w ? w->tol : 1e-8; op.max_iter = w ? w->max_iter : 10;
    op.active = active; op.iters = iters; op.outer_iters = oit; op.n_fail = n_fail;
    Ptrs<2, 3> ptrs; ptrs.in[0] = deps; ptrs.in[1] = sin; ptrs.out[0] = sig; ptrs.out[1] = tan; ptrs.out[2] = sout;
    return launch_points<WrappedCrystalOp<KIND>, 64, 32>(op, ptrs, N, layout, st, "mm_wrapped_crystal");
}

template <int KIND>
int launch_wrapped_elastic(const MMLinearElastic* p, const double* eps, double* sig, double* tan, uint8_t* oit, int32_t* n_fail, const MMNewtonOpts* w,
                           int64_t N, int layout, cudaStream_t st) {
    WrappedElasticOp<KIND> op;
    op.k = make_elastic_k(p->E, p->nu); op.tol = w ? w->tol : 1e-8; op.max_iter = w ? w->max_iter : 10; op.outer_iters = oit; op.n_fail = n_fail;
    Ptrs<1, 2> ptrs; ptrs.in[0] = eps; ptrs.out[0] = sig; ptrs.out[1] = tan;
    return launch_points<WrappedElasticOp<KIND>, 256, 128>(op, ptrs, N, layout, st, "mm_wrapped_linear_elastic");
}
// ---------------------------------------------------------------------------------------------------------
// Lane-persistent kernel for the stress-restricted kinds around J2 (the state machine of mm_wrap_lp.cuh).  Each warp owns a contiguous
// chunk of points; a lane carries one point through TRIAL / NEWTON / OUTER / FINAL and takes the next point of the chunk when it is
// through, so a warp no longer runs the maximum of its lanes' 2-6 inner solves x 0-3 Newton passes.  Every section is entered on a
// warp vote; the two sections that only some lanes need per trip (OUTER, FINAL) wait until `th_outer` / `th_final` lanes want them
// or nothing else can advance.  Global accesses are per lane (a lane's point is wherever the queue was when it finished), but the lanes
// of a warp stay within a window of a few dozen points, so L1/L2 see whole sectors; HBM traffic is the algorithmic figure.
struct WrapKnobs { int variant, th_outer, th_final, chunk; };
static WrapKnobs& wrap_knobs() {
    static WrapKnobs k = [] {
        WrapKnobs v{-1, 0, 0, 0};
        if (const char* e = getenv("MM_WRAPPED_J2_VARIANT")) v.variant = atoi(e);
        if (const char* e = getenv("MM_WRAPPED_J2_TH_OUTER")) v.th_outer = atoi(e);
        if (const char* e = getenv("MM_WRAPPED_J2_TH_FINAL")) v.th_final = atoi(e);
        if (const char* e = getenv("MM_WRAPPED_J2_CHUNK")) v.chunk = atoi(e);
        return v;
    }();
    return k;
}
int* wrap_tuning_slot(const char* name) {
    WrapKnobs& k = wrap_knobs();
    if (!strcmp(name, "wrapped_j2_variant")) return &k.variant;          // 0 = one thread per point, 1 (default) = lane-persistent
    if (!strcmp(name, "wrapped_j2_th_outer")) return &k.th_outer;
    if (!strcmp(name, "wrapped_j2_th_final")) return &k.th_final;
    if (!strcmp(name, "wrapped_j2_chunk")) return &k.chunk;
    return nullptr;
}

template <int KIND, int LAYOUT> struct WrapLaneAcc {
    const Ptrs<2, 3>& p;
    int64_t N, i;
    __device__ __forceinline__ WrapLaneAcc(const Ptrs<2, 3>& p_, int64_t N_, int64_t i_) : p(p_), N(N_), i(i_) {}
    template <int A> __device__ __forceinline__ double in(int c) const {
        return LAYOUT == 0 ? p.in[A][(int64_t)c * N + i] : p.in[A][i * WrappedPlasticOp<KIND>::in_nc(A) + c];
    }
    template <int A> __device__ __forceinline__ void out(int c, double v) const {
        if (LAYOUT == 0) p.out[A][(int64_t)c * N + i] = v; else p.out[A][i * WrappedPlasticOp<KIND>::out_nc(A) + c] = v;
    }
    template <int A> __device__ __forceinline__ bool has_out() const { return p.out[A] != nullptr; }
};

#ifndef MM_WRAPLP_BS
#define MM_WRAPLP_BS 128
#endif
#ifndef MM_WRAPLP_MINBLOCKS
#define MM_WRAPLP_MINBLOCKS 3
#endif
#ifndef MM_WRAPLP_TH_OUTER
#define MM_WRAPLP_TH_OUTER 8
#endif
#ifndef MM_WRAPLP_TH_FINAL
#define MM_WRAPLP_TH_FINAL 32     // PlaneStress / ShellStress; UniaxialStress (1 reduced component, 1 tangent entry: a cheap FINAL section) uses MM_WRAPLP_TH_FINAL_UNIAXIAL
#endif
#ifndef MM_WRAPLP_TH_FINAL_UNIAXIAL
#define MM_WRAPLP_TH_FINAL_UNIAXIAL 4
#endif
template <int KIND, int LAYOUT, int BS, bool EVAL_TRIPS>
__global__ void __launch_bounds__(BS, MM_WRAPLP_MINBLOCKS)
wrapped_j2_lp_kernel(const PlasticK k, const double tol, const int max_iter, const Ptrs<2, 3> p, uint8_t* flags, uint16_t* iters, uint8_t* outer_iters, int32_t* n_fail,
                     const int64_t N, const int chunk, const int th_outer, const int th_final) {
    extern __shared__ double sm_scratch[];
    const Scratch scr{sm_scratch + threadIdx.x, BS};
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t warp_id = (int64_t)blockIdx.x * (BS / 32) + (threadIdx.x >> 5);
    int64_t next = warp_id * chunk;
    const int64_t end = (next + chunk < N) ? (next + chunk) : N;
    J2Lane L;
    L.phase = J2P_IDLE;
    int64_t my = 0;
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, L.phase == J2P_IDLE);
        if (idle != 0u && next < end) {                       // refill idle lanes with the next points of the chunk
            if (L.phase == J2P_IDLE) {
                const int64_t cand = next + __popc(idle & ((1u << lane) - 1u));
                if (cand < end) {
                    my = cand;
                    WrapLaneAcc<KIND, LAYOUT> acc(p, N, my);
                    j2lp_load<KIND>(acc, scr, L);
                }
            }
            next += __popc(idle);
        }
        if (__ballot_sync(FULL, L.phase != J2P_IDLE) == 0u) break;
        if (L.phase == J2P_TRIAL) j2lp_trial(k, scr, L);
        if (L.phase == J2P_NEWTON) j2lp_pass(k, scr, L);
        if (EVAL_TRIPS) {                                     // one trip = one whole inner solve: the passes of a yielded evaluation are 3 almost always
            while (__ballot_sync(FULL, L.phase == J2P_NEWTON) != 0u) { if (L.phase == J2P_NEWTON) j2lp_pass(k, scr, L); }
        }
        const unsigned iterating = __ballot_sync(FULL, L.phase == J2P_NEWTON);
        const unsigned want_outer = __ballot_sync(FULL, L.phase == J2P_OUTER);
        if (want_outer != 0u && (__popc(want_outer) >= th_outer || iterating == 0u)) {
            if (L.phase == J2P_OUTER) j2lp_outer<KIND>(k, scr, L, tol, max_iter);
        }
        const unsigned want_final = __ballot_sync(FULL, L.phase == J2P_FINAL);
        const unsigned advancing = __ballot_sync(FULL, L.phase == J2P_TRIAL || L.phase == J2P_NEWTON || L.phase == J2P_OUTER);
        if (want_final != 0u && (__popc(want_final) >= th_final || advancing == 0u || next >= end)) {
            if (L.phase == J2P_FINAL) {
                WrapLaneAcc<KIND, LAYOUT> acc(p, N, my);
                if (!j2lp_final<KIND>(k, scr, L, acc)) L.flag |= 0x40;
                if (flags) flags[my] = (uint8_t)L.flag;
                if (iters) iters[my] = (uint16_t)(L.iters > 65535 ? 65535 : L.iters);
                if (outer_iters) outer_iters[my] = (uint8_t)L.oit;
                if ((L.flag & 0xC0) && n_fail) atomicAdd(n_fail, 1);
                L.phase = J2P_IDLE;
            }
        }
    }
}
template <int KIND, int LAYOUT>
int launch_wrapped_j2_lp(const PlasticK& k, double tol, int max_iter, const Ptrs<2, 3>& ptrs, uint8_t* flags, uint16_t* iters, uint8_t* oit, int32_t* n_fail, int64_t N,
                         cudaStream_t st) {
    constexpr int BS = MM_WRAPLP_BS;
    constexpr size_t smem = (size_t)J2S_SLOTS * BS * sizeof(double);
    static_assert(smem <= 48 * 1024, "scratch columns beyond 48 KB need the shared-memory opt-in");
    const WrapKnobs& kn = wrap_knobs();
    // chunk of points per warp: long enough that the drain at its end (lanes running dry) is a small share, short enough for many waves
    const int64_t resident_warps = (int64_t)sm_count() * MM_WRAPLP_MINBLOCKS * (BS / 32);
    int64_t chunk = N / (resident_warps * 8);
    if (chunk > 2048) chunk = 2048;
    if (chunk < 128) chunk = 128;
    chunk = (chunk + 31) / 32 * 32;
    if (kn.chunk > 0) chunk = kn.chunk;
    const int64_t warps = (N + chunk - 1) / chunk;
    const int64_t blocks = (warps + (BS / 32) - 1) / (BS / 32);
    if (blocks > 2147483647LL) return set_error(-1, "mm_wrapped_plastic: N too large for one launch");
    const int th_outer = kn.th_outer > 0 ? kn.th_outer : MM_WRAPLP_TH_OUTER, th_final = kn.th_final > 0 ? kn.th_final : (KIND == WRAP_UNIAXIAL_STRESS ? MM_WRAPLP_TH_FINAL_UNIAXIAL : MM_WRAPLP_TH_FINAL);
    if (kn.variant == 2) wrapped_j2_lp_kernel<KIND, LAYOUT, BS, false><<<(unsigned)blocks, BS, smem, st>>>(k, tol, max_iter, ptrs, flags, iters, oit, n_fail, N, (int)chunk, th_outer, th_final);
    else                 wrapped_j2_lp_kernel<KIND, LAYOUT, BS, true><<<(unsigned)blocks, BS, smem, st>>>(k, tol, max_iter, ptrs, flags, iters, oit, n_fail, N, (int)chunk, 1, th_final);
    return check_launch("mm_wrapped_plastic (lane-persistent kernel)");
}

template <int KIND>
int launch_wrapped_plastic(const MMPlastic* p, const double* eps, const double* sin, double* sig, double* tan, double* sout, uint8_t* flags, uint16_t* iters,
                           uint8_t* oit, int32_t* n_fail, const MMNewtonOpts* o, const MMNewtonOpts* w, int64_t N, int layout, cudaStream_t st) {
    WrappedPlasticOp<KIND> op;
    op.k = make_plastic_k(p->E, p->nu, p->sigma_y, p->H, p->r, p->kappa_inf, p->alpha_inf, o ? o->tol : 1e-8, o ? o->max_iter : 1000);
    op.tol = w ? w->tol : 1e-8; op.max_iter = w ? w->max_iter : 10;
    op.flags = flags; op.iters = iters; op.outer_iters = oit; op.n_fail = n_fail;
    Ptrs<2, 3> ptrs; ptrs.in[0] = eps; ptrs.in[1] = sin; ptrs.out[0] = sig; ptrs.out[1] = tan; ptrs.out[2] = sout;
    if constexpr (KIND == WRAP_UNIAXIAL_STRESS || KIND == WRAP_PLANE_STRESS || KIND == WRAP_SHELL_STRESS) {
        if (wrap_knobs().variant != 0 && N > 0 && (layout == 0 || layout == 1)) {
            return layout == 0 ? launch_wrapped_j2_lp<KIND, 0>(op.k, op.tol, op.max_iter, ptrs, flags, iters, oit, n_fail, N, st)
                               : launch_wrapped_j2_lp<KIND, 1>(op.k, op.tol, op.max_iter, ptrs, flags, iters, oit, n_fail, N, st);
        }
    }
    return launch_points<WrappedPlasticOp<KIND>, 128, 64>(op, ptrs, N, layout, st, "mm_wrapped_plastic");
}

}  // namespace mm

using namespace mm;

extern "C" {

int mm_wrap_ncomp(int dim_kind) {
    if (dim_kind < MM_DIM_UNIAXIAL_STRAIN || dim_kind > MM_DIM_SHELL_STRESS) return set_error(MM_ERR_ARG, "mm_wrap_ncomp: unknown dim_kind %d", dim_kind);
    return wrap_nc(dim_kind);
}

int mm_wrapped_linear_elastic(int dim_kind, const MMLinearElastic* p, const double* eps_red, double* sig_red, double* tan_red, uint8_t* outer_iters,
                              int32_t* n_fail, const MMNewtonOpts* wrapper_opts, int64_t N, int layout, void* stream) {
    if (!p || N < 0 || (N > 0 && (!eps_red || !sig_red))) return set_error(MM_ERR_ARG, "mm_wrapped_linear_elastic: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    switch (dim_kind) {
        case MM_DIM_UNIAXIAL_STRAIN: return launch_wrapped_elastic<WRAP_UNIAXIAL_STRAIN>(p, eps_red, sig_red, tan_red, outer_iters, n_fail, wrapper_opts, N, layout, st);
        case MM_DIM_PLANE_STRAIN: return launch_wrapped_elastic<WRAP_PLANE_STRAIN>(p, eps_red, sig_red, tan_red, outer_iters, n_fail, wrapper_opts, N, layout, st);
        case MM_DIM_UNIAXIAL_STRESS: return launch_wrapped_elastic<WRAP_UNIAXIAL_STRESS>(p, eps_red, sig_red, tan_red, outer_iters, n_fail, wrapper_opts, N, layout, st);
        case MM_DIM_PLANE_STRESS: return launch_wrapped_elastic<WRAP_PLANE_STRESS>(p, eps_red, sig_red, tan_red, outer_iters, n_fail, wrapper_opts, N, layout, st);
        case MM_DIM_SHELL_STRESS: return launch_wrapped_elastic<WRAP_SHELL_STRESS>(p, eps_red, sig_red, tan_red, outer_iters, n_fail, wrapper_opts, N, layout, st);
    }
    return set_error(MM_ERR_ARG, "mm_wrapped_linear_elastic: unknown dim_kind %d", dim_kind);
}

int mm_wrapped_transversely_isotropic(int dim_kind, const MMTransverselyIsotropic* p, const double* eps_red, const double* a3, double* sig_red, double* tan_red,
                                      uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* wrapper_opts, int64_t N, int layout, void* stream) {
    if (!p || N < 0 || (N > 0 && (!eps_red || !a3 || !sig_red))) return set_error(MM_ERR_ARG, "mm_wrapped_transversely_isotropic: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
#define MM_WT(K) return launch_wrapped_transviso<K>(p, eps_red, a3, sig_red, tan_red, outer_iters, n_fail, wrapper_opts, N, layout, st)
    switch (dim_kind) {
        case MM_DIM_UNIAXIAL_STRAIN: MM_WT(WRAP_UNIAXIAL_STRAIN);
        case MM_DIM_PLANE_STRAIN: MM_WT(WRAP_PLANE_STRAIN);
        case MM_DIM_UNIAXIAL_STRESS: MM_WT(WRAP_UNIAXIAL_STRESS);
        case MM_DIM_PLANE_STRESS: MM_WT(WRAP_PLANE_STRESS);
        case MM_DIM_SHELL_STRESS: MM_WT(WRAP_SHELL_STRESS);
    }
#undef MM_WT
    return set_error(MM_ERR_ARG, "mm_wrapped_transversely_isotropic: unknown dim_kind %d", dim_kind);
}

int mm_wrapped_crystal(int dim_kind, const MMCrystal* p, const double* deps_red, double dt, const double* state_in, double* sig_red, double* tan_red,
                       double* state_out, uint16_t* active, uint16_t* iters, uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* opts,
                       const MMNewtonOpts* wrapper_opts, int64_t N, int layout, void* stream) {
    if (!p || N < 0 || (N > 0 && (!deps_red || !state_in || !sig_red || !state_out))) return set_error(MM_ERR_ARG, "mm_wrapped_crystal: bad argument");
    if (p->nslip != MM_NSLIP) return set_error(MM_ERR_ARG, "mm_wrapped_crystal: nslip must be %d (got %d)", MM_NSLIP, p->nslip);
    CrystalK k;
    make_crystal_k(k, p->E, p->nu, p->tau_y, p->H_kin, p->alpha_inf, p->t_star, p->sigma_c, p->m, p->MS, p->H, dt, opts ? opts->tol : 1e-8,
                   opts ? opts->max_iter : 100);
    cudaStream_t st = (cudaStream_t)stream;
#define MM_WC(K) return launch_wrapped_crystal<K>(k, deps_red, state_in, sig_red, tan_red, state_out, active, iters, outer_iters, n_fail, wrapper_opts, N, layout, st)
    switch (dim_kind) {
        case MM_DIM_UNIAXIAL_STRAIN: MM_WC(WRAP_UNIAXIAL_STRAIN);
        case MM_DIM_PLANE_STRAIN: MM_WC(WRAP_PLANE_STRAIN);
        case MM_DIM_UNIAXIAL_STRESS: MM_WC(WRAP_UNIAXIAL_STRESS);
        case MM_DIM_PLANE_STRESS: MM_WC(WRAP_PLANE_STRESS);
        case MM_DIM_SHELL_STRESS: MM_WC(WRAP_SHELL_STRESS);
    }
#undef MM_WC
    return set_error(MM_ERR_ARG, "mm_wrapped_crystal: unknown dim_kind %d", dim_kind);
}

int mm_wrapped_plastic(int dim_kind, const MMPlastic* p, const double* eps_red, const double* state_in, double* sig_red, double* tan_red, double* state_out,
                       uint8_t* flags, uint16_t* iters, uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* opts, const MMNewtonOpts* wrapper_opts,
                       int64_t N, int layout, void* stream) {
    if (!p || N < 0 || (N > 0 && (!eps_red || !state_in || !sig_red || !state_out))) return set_error(MM_ERR_ARG, "mm_wrapped_plastic: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
#define MM_WP(K) return launch_wrapped_plastic<K>(p, eps_red, state_in, sig_red, tan_red, state_out, flags, iters, outer_iters, n_fail, opts, wrapper_opts, N, layout, st)
    switch (dim_kind) {
        case MM_DIM_UNIAXIAL_STRAIN: MM_WP(WRAP_UNIAXIAL_STRAIN);
        case MM_DIM_PLANE_STRAIN: MM_WP(WRAP_PLANE_STRAIN);
        case MM_DIM_UNIAXIAL_STRESS: MM_WP(WRAP_UNIAXIAL_STRESS);
        case MM_DIM_PLANE_STRESS: MM_WP(WRAP_PLANE_STRESS);
        case MM_DIM_SHELL_STRESS: MM_WP(WRAP_SHELL_STRESS);
    }
#undef MM_WP
    return set_error(MM_ERR_ARG, "mm_wrapped_plastic: unknown dim_kind %d", dim_kind);
}

}  // extern "C"
