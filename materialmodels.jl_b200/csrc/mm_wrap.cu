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
struct WrapKnobs { int variant, th_outer, th_final, chunk, crystal_variant, crystal_slab; };
static WrapKnobs& wrap_knobs() {
    static WrapKnobs k = [] {
        WrapKnobs v{-1, 0, 0, 0, -1, 0};
        if (const char* e = getenv("MM_WRAPPED_J2_VARIANT")) v.variant = atoi(e);
        if (const char* e = getenv("MM_WRAPPED_J2_TH_OUTER")) v.th_outer = atoi(e);
        if (const char* e = getenv("MM_WRAPPED_J2_TH_FINAL")) v.th_final = atoi(e);
        if (const char* e = getenv("MM_WRAPPED_J2_CHUNK")) v.chunk = atoi(e);
        if (const char* e = getenv("MM_WRAPPED_CRYSTAL_VARIANT")) v.crystal_variant = atoi(e);
        if (const char* e = getenv("MM_WRAPPED_CRYSTAL_SLAB")) v.crystal_slab = atoi(e);
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
    if (!strcmp(name, "wrapped_crystal_variant")) return &k.crystal_variant;   // 0 = one thread per point, 1 (default) = rounds of the two-kernel crystal path
    if (!strcmp(name, "wrapped_crystal_slab")) return &k.crystal_slab;         // points per workspace slab (default 2^21)
    return nullptr;
}

// ---------------------------------------------------------------------------------------------------------
// Stress-restricted kinds around CrystalViscoPlastic as ROUNDS of the tuned two-kernel crystal path (mm_crystal.cu): every outer
// iteration of wrappers.jl:60-77 is a full inner crystal solve (5-30 Newton passes, wildly different between neighbours) plus its
// tangent — exactly what crystal_newton_kernel (lane-persistent) + crystal_finish_kernel do well and a one-thread-per-point wrapper
// does badly (255 registers with spills, a warp runs its slowest lane's solve on every outer iteration: 5.3-6.3e7 updates/s).  Per round r:
//   1. the crystal kernels on the BATCH of still-unconverged points: a device-side index list idx[r][0 .. cnt[r-1]); their inputs
//      (3D strain, old state) are workspace arrays indexed by point — nothing is moved between rounds —, their outputs (σ, ∂σ∂ε, new
//      state) are packed by list position (scattered 8-byte stores of 84 doubles per point cost 0.6 ms per round whatever the list length);
//   2. wrapped_crystal_outer_kernel: out-of-plane stress norm; converged -> Schur-complement tangent + all outputs of the point;
//      otherwise one outer Newton update of the point's 3D strain in place, and the point is appended to the next round's list
//      (warp-aggregated atomic on cnt[r]; the order of a list does not matter — points are independent).
// max_iter rounds are always enqueued (no host synchronisation: the call stays stream-ordered and graph-capturable); a round whose
// list is empty costs four empty launches.  Same iterates as the one-thread-per-point path: same inner solves from the same 3D
// strains, wrap_outer_step / wrap_schur_out are the expressions of wrapped_core.  Large batches go through the workspace in slabs.
int crystal_structured_batch(const CrystalK& k, uint16_t* active, uint16_t* iters, const Ptrs<2, 3>& ptrs, int64_t N, const int* cnt, const int* idx, cudaStream_t st);   // mm_crystal.cu
bool crystal_force_general();                                                                                                                                         // mm_crystal.cu

struct WrapXtalWs {
    double* eps;           // [6][cap]  by point: 3D strain increment (updated in place by the outer step)
    double* state;         // [42][cap] by point: old state (copied once from the caller's array, which may be AoS / aliased with state_out)
    double* sig;           // [6][cap]  by list position: σ of the round's inner solve
    double* tan;           // [36][cap] by list position: ∂σ∂ε
    double* stout;         // [42][cap] by list position: new state
    int* idx[2];           // [cap]     the round's list of points (ping-pong)
    uint16_t* act;         // [cap]     by list position: inner active mask / status
    uint16_t* its;         // [cap]     by list position: inner Newton count
    int* cnt;              // [max_iter + 1] list lengths per round
};
struct WrapXtalIO {        // the caller's arrays (either layout) and the slab
    const double* eps_red; const double* state_in;
    double* sig_red; double* tan_red; double* state_out;
    uint16_t* active; uint16_t* iters; uint8_t* outer_iters; int32_t* n_fail;
    int64_t N, i0;         // all points (= SoA stride of the caller's arrays), first point of the slab
};
template <int LAYOUT> __device__ __forceinline__ int64_t wx_at(int64_t N, int64_t i, int nc, int c) { return LAYOUT == 0 ? (int64_t)c * N + i : i * nc + c; }

template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(256) wrapped_crystal_init_kernel(const WrapXtalWs w, const WrapXtalIO io, const int n, const int cap) {
    constexpr int d = wrap_dim(KIND), nc = wrap_nc(KIND);
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j == 0) w.cnt[0] = n;
    if (j >= n) return;
    const int64_t i = io.i0 + j;
    double e[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int r = 0; r < nc; ++r) e[wrap_slot(d, r)] = io.eps_red[wx_at<LAYOUT>(io.N, i, nc, r)];       // increase_dim (wrappers.jl:24-27)
#pragma unroll
    for (int c = 0; c < 6; ++c) w.eps[(int64_t)c * cap + j] = e[c];
#pragma unroll
    for (int c0 = 0; c0 < 42; c0 += 14) {
        double v[14];
#pragma unroll
        for (int c = 0; c < 14; ++c) v[c] = io.state_in[wx_at<LAYOUT>(io.N, i, 42, c0 + c)];
#pragma unroll
        for (int c = 0; c < 14; ++c) w.state[(int64_t)(c0 + c) * cap + j] = v[c];
    }
    w.idx[0][j] = j;
}

template <int KIND, int LAYOUT> struct WrapXtalOutAcc {        // reduced outputs of one point into the caller's arrays
    const WrapXtalIO& io; int64_t i;
    template <int A> __device__ __forceinline__ void out(int c, double v) const {
        constexpr int nc = wrap_nc(KIND);
        if (A == 0) io.sig_red[wx_at<LAYOUT>(io.N, i, nc, c)] = v;
        else if (A == 1) io.tan_red[wx_at<LAYOUT>(io.N, i, nc * nc, c)] = v;
    }
    template <int A> __device__ __forceinline__ bool has_out() const { return A == 1 ? io.tan_red != nullptr : true; }
};
struct WrapXtalTangent {
    const double* tan; int64_t cap; int j;
    __device__ __forceinline__ double operator()(int I, int J) const { return tan[(int64_t)(I + 6 * J) * cap + j]; }
};

template <int KIND, int LAYOUT>
__global__ void __launch_bounds__(128) wrapped_crystal_outer_kernel(const WrapXtalWs w, const WrapXtalIO io, const double tol, const int it, const int max_iter, const int cap) {
    constexpr int d = wrap_dim(KIND), nc = wrap_nc(KIND);
    // (a run-time index into the kernel parameter's idx[2] would move the whole parameter block to local memory)
    const int* idx_in = ((it - 1) & 1) ? w.idx[1] : w.idx[0];
    int* idx_out = ((it - 1) & 1) ? w.idx[0] : w.idx[1];
    const int n = w.cnt[it - 1];
    const int j = blockIdx.x * 128 + threadIdx.x;
    bool again = false;
    int li = 0;
    if (j < n) {
        li = idx_in[j];
        const int64_t i = io.i0 + li;
        const unsigned act = w.act[j];
        int flag = (act & MM_ACTIVE_FAILED) ? 0x80 : 0, oit = it - 1;
        bool done = false;
        double sig[6], eps[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) { sig[c] = w.sig[(int64_t)c * cap + j]; eps[c] = w.eps[(int64_t)c * cap + li]; }
        const WrapXtalTangent T{w.tan, cap, j};
        if (!(flag & 0x80)) {
            const int r = wrap_outer_step<KIND>(T, sig, eps, tol);
            if (r == 0) done = true;
            else if (r == 2) flag |= 0x40;
            else { oit = it; if (it >= max_iter) flag |= 0x40; else again = true; }      // "Could not find plane stress state" (wrappers.jl:78)
        }
        if (again) {
#pragma unroll
            for (int a = 0; a < wrap_nzero(KIND); ++a) { const int c = slot_of_mandel(wrap_zero(KIND, a)); w.eps[(int64_t)c * cap + li] = eps[c]; }
        } else {                                                                         // this point is through: all its outputs
            WrapXtalOutAcc<KIND, LAYOUT> acc{io, i};
#pragma unroll
            for (int r = 0; r < nc; ++r) acc.template out<0>(r, sig[wrap_slot(d, r)]);
            if (done) { if (!wrap_schur_out<KIND>(T, acc)) flag |= 0x40; }
            else if (io.tan_red) {
#pragma unroll
                for (int c = 0; c < nc * nc; ++c) acc.template out<1>(c, 0.0);
            }
#pragma unroll
            for (int c0 = 0; c0 < 42; c0 += 14) {
                double v[14];
#pragma unroll
                for (int c = 0; c < 14; ++c) v[c] = w.stout[(int64_t)(c0 + c) * cap + j];
#pragma unroll
                for (int c = 0; c < 14; ++c) io.state_out[wx_at<LAYOUT>(io.N, i, 42, c0 + c)] = v[c];
            }
            unsigned mask = act;
            if (flag & 0x40) mask |= MM_ACTIVE_WRAPPER_FAILED;
            if (io.active) io.active[i] = (uint16_t)mask;
            if (io.iters) io.iters[i] = w.its[j];
            if (io.outer_iters) io.outer_iters[i] = (uint8_t)oit;
            if ((flag & 0xC0) && io.n_fail) atomicAdd(io.n_fail, 1);
        }
    }
    // append the points that go on to the next round's list (one atomic per warp)
    const unsigned m = __ballot_sync(0xffffffffu, again);
    if (m != 0u) {
        const int lane = threadIdx.x & 31, leader = __ffs((int)m) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&w.cnt[it], __popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (again) idx_out[base + __popc(m & ((1u << lane) - 1u))] = li;
    }
}

template <int KIND, int LAYOUT>
int launch_wrapped_crystal_rounds(const CrystalK& k, const WrapXtalIO& io_all, const MMNewtonOpts* wopts, int64_t N, cudaStream_t st) {
    const double tol = wopts ? wopts->tol : 1e-8;
    const int max_iter = wopts ? wopts->max_iter : 10;
    if (max_iter < 1 || max_iter > 64) return -100;                  // nothing to iterate / an unusual cap (4 launches per round are enqueued whatever
                                                                     // converges): the caller falls back to the one-thread-per-point kernel
    int64_t slab = wrap_knobs().crystal_slab > 0 ? wrap_knobs().crystal_slab : (1 << 21);
    if (slab > N) slab = N;
    const int cap = (int)((slab + 1) & ~(int64_t)1);                 // even: every component column of the workspace stays 16-byte aligned
    const size_t nd = (size_t)(6 + 42 + 6 + 36 + 42) * cap;
    const size_t bytes = nd * sizeof(double) + (size_t)2 * cap * sizeof(int) + (size_t)2 * cap * sizeof(uint16_t) + (size_t)(max_iter + 2) * sizeof(int) + 64;
    char* raw = nullptr;
    cudaError_t e = ws_alloc_async((void**)&raw, bytes, st);
    if (e != cudaSuccess) return set_error(MM_ERR_CUDA, "mm_wrapped_crystal: workspace of %zu bytes: %s", bytes, cudaGetErrorString(e));
    WrapXtalWs w;
    double* dp = reinterpret_cast<double*>(raw);
    w.eps = dp; dp += (size_t)6 * cap;
    w.state = dp; dp += (size_t)42 * cap;
    w.sig = dp; dp += (size_t)6 * cap;
    w.tan = dp; dp += (size_t)36 * cap;
    w.stout = dp; dp += (size_t)42 * cap;
    int* ip = reinterpret_cast<int*>(dp);
    w.idx[0] = ip; ip += cap;
    w.idx[1] = ip; ip += cap;
    w.cnt = ip; ip += max_iter + 1 + ((max_iter + 1) & 1);
    w.act = reinterpret_cast<uint16_t*>(ip);
    w.its = w.act + cap;
    int rc = MM_OK;
    for (int64_t i0 = 0; i0 < N && rc == MM_OK; i0 += slab) {
        const int n = (int)((N - i0 < slab) ? (N - i0) : slab);
        WrapXtalIO io = io_all;
        io.i0 = i0;
        cudaMemsetAsync(w.cnt, 0, (size_t)(max_iter + 1) * sizeof(int), st);
        wrapped_crystal_init_kernel<KIND, LAYOUT><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, io, n, cap);
        if ((rc = check_launch("mm_wrapped_crystal (init)")) != MM_OK) break;
        for (int it = 1; it <= max_iter; ++it) {
            Ptrs<2, 3> ptrs; ptrs.in[0] = w.eps; ptrs.in[1] = w.state; ptrs.out[0] = w.sig; ptrs.out[1] = w.tan; ptrs.out[2] = w.stout;
            if ((rc = crystal_structured_batch(k, w.act, w.its, ptrs, cap, w.cnt + (it - 1), w.idx[(it - 1) & 1], st)) != MM_OK) break;
            wrapped_crystal_outer_kernel<KIND, LAYOUT><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(w, io, tol, it, max_iter, cap);
            if ((rc = check_launch("mm_wrapped_crystal (outer step)")) != MM_OK) break;
        }
    }
    cudaFreeAsync(raw, st);
    return rc;
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
    // stress-restricted kinds with the reference constructor's hardening matrix: rounds of the tuned two-kernel crystal path
    if (dim_kind >= MM_DIM_UNIAXIAL_STRESS && dim_kind <= MM_DIM_SHELL_STRESS && N > 0 && (layout == 0 || layout == 1) && k.h_struct && !crystal_force_general() &&
        wrap_knobs().crystal_variant != 0) {
        const WrapXtalIO io{deps_red, state_in, sig_red, tan_red, state_out, active, iters, outer_iters, n_fail, N, 0};
        int rc = -100;
#define MM_WR(K) rc = (layout == 0) ? launch_wrapped_crystal_rounds<K, 0>(k, io, wrapper_opts, N, st) : launch_wrapped_crystal_rounds<K, 1>(k, io, wrapper_opts, N, st)
        if (dim_kind == MM_DIM_UNIAXIAL_STRESS) MM_WR(WRAP_UNIAXIAL_STRESS);
        else if (dim_kind == MM_DIM_PLANE_STRESS) MM_WR(WRAP_PLANE_STRESS);
        else MM_WR(WRAP_SHELL_STRESS);
#undef MM_WR
        if (rc != -100) return rc;
    }
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
