// TEST-ONLY: instantiates the kernels' per-point math (csrc/mm_math.cuh, mm_crystal_math.cuh) with a
// host accessor so the algebra can be compared with the oracle on a CPU-only box.  Never shipped,
// never loaded by the product; the GPU parity tests (-m gpu) are the real gate.
#include <cstdint>
#include <cstring>
#define MM_CRYSTAL_STATS 1
namespace mm { void mm_crystal_stats_pass(unsigned nz, unsigned cand); }
static inline void trace_put(uint8_t v);
#include "../../include/mmb200.h"
#include "../../materialmodels.jl_b200/csrc/mm_math.cuh"
#include "../../materialmodels.jl_b200/csrc/mm_crystal_math.cuh"
#include "../../materialmodels.jl_b200/csrc/mm_wrap_math.cuh"
#include "../../materialmodels.jl_b200/csrc/mm_wrap_lp.cuh"

namespace {
template <int NIN, int NOUT> struct HostAcc {   // SoA host arrays
    const double* inp[NIN]; double* outp[NOUT]; int64_t N, i;
    template <int A> double in(int c) const { return inp[A][(int64_t)c * N + i]; }
    template <int A> void out(int c, double v) const { outp[A][(int64_t)c * N + i] = v; }
    template <int A> bool has_out() const { return outp[A] != nullptr; }
};
}

extern "C" {
int twin_linear_elastic(const MMLinearElastic* p, const double* eps, double* sig, double* tan, int64_t N) {
    mm::ElasticK k = mm::make_elastic_k(p->E, p->nu);
    for (int64_t i = 0; i < N; ++i) { HostAcc<1, 2> a{{eps}, {sig, tan}, N, i}; mm::linear_elastic_point(k, a); }
    return 0;
}
int twin_plastic(const MMPlastic* p, const double* eps, const double* sin, double* sig, double* tan, double* sout, uint8_t* flags, uint16_t* iters,
                 double tol, int max_iter, int64_t N) {
    mm::PlasticK k = mm::make_plastic_k(p->E, p->nu, p->sigma_y, p->H, p->r, p->kappa_inf, p->alpha_inf, tol, max_iter);
    for (int64_t i = 0; i < N; ++i) {
        HostAcc<2, 3> a{{eps, sin}, {sig, tan, sout}, N, i};
        int it = 0; int f = mm::plastic_point(k, a, &it);
        flags[i] = (uint8_t)f; iters[i] = (uint16_t)it;
    }
    return 0;
}
// the factored tangent {gξ, u, n} (parts[13][N]) of the same update: what the compact return ships for a yielded point
int twin_plastic_parts(const MMPlastic* p, const double* eps, const double* sin, double* sig, double* parts, double* sout, uint8_t* flags, uint16_t* iters,
                       double tol, int max_iter, int64_t N) {
    mm::PlasticK k = mm::make_plastic_k(p->E, p->nu, p->sigma_y, p->H, p->r, p->kappa_inf, p->alpha_inf, tol, max_iter);
    for (int64_t i = 0; i < N; ++i) {
        HostAcc<2, 3> a{{eps, sin}, {sig, nullptr, sout}, N, i};
        int it = 0; mm::PlasticTangentParts tp;
        int f = mm::plastic_point_parts(k, a, &it, tp);
        parts[i] = tp.gxi;
        for (int c = 0; c < 6; ++c) { parts[(int64_t)(1 + c) * N + i] = tp.u[c]; parts[(int64_t)(7 + c) * N + i] = tp.n[c]; }
        flags[i] = (uint8_t)f; iters[i] = (uint16_t)it;
    }
    return 0;
}
int twin_hyper(int model, const MMHyper* p, const double* strain, int kind, double* S, double* T, int64_t N) {
    mm::HyperK k{p->lambda, p->mu, p->c2, p->c3};
    for (int64_t i = 0; i < N; ++i) {
        HostAcc<1, 2> a{{strain}, {S, T}, N, i};
        if (kind == 1) { if (model == 0) mm::hyper_point<0, 1>(k, a); else if (model == 1) mm::hyper_point<1, 1>(k, a); else mm::hyper_point<2, 1>(k, a); }
        else           { if (model == 0) mm::hyper_point<0, 0>(k, a); else if (model == 1) mm::hyper_point<1, 0>(k, a); else mm::hyper_point<2, 0>(k, a); }
    }
    return 0;
}
int twin_xu(const MMXuNeedleman* p, const double* jump, double* T, double* H, int dim, int64_t N) {
    mm::XuK k = mm::make_xu_k(p->Phi_n, p->Phi_t, p->delta_n, p->delta_t, p->Delta_n_star);
    for (int64_t i = 0; i < N; ++i) { HostAcc<1, 2> a{{jump}, {T, H}, N, i}; if (dim == 3) mm::xu_point<3>(k, a); else mm::xu_point<2>(k, a); }
    return 0;
}
struct HostScratch { double v[MM_CRYSTAL_SCRATCH + 8]; double& operator()(int j) { return v[j]; } };     // 72 slots for the phased pass

// mode 0: general pivoted path; mode 1: structured fast path (falls back like the kernel does); mode 2: the same through the sparse pass; mode 3: through the phased pass (q = 0; else as mode 2); mode 4: through the active-set pass
int twin_crystal_mode(const MMCrystal* p, const double* deps, double dt, const double* sin, double* sig, double* tan, double* sout, uint16_t* active,
                      uint16_t* iters, double tol, int max_iter, int64_t N, int mode, int64_t* n_fallback) {
    mm::CrystalK k;
    mm::make_crystal_k(k, p->E, p->nu, p->tau_y, p->H_kin, p->alpha_inf, p->t_star, p->sigma_c, p->m, p->MS, p->H, dt, tol, max_iter);
    int64_t nfb = 0;
    for (int64_t i = 0; i < N; ++i) {
        HostAcc<2, 3> a{{deps, sin}, {sig, tan, sout}, N, i};
        int it = 0; unsigned m;
        trace_put(255);
        if (mode == 0 || !k.h_struct) m = mm::crystal_point(k, a, &it);
        else {
            HostScratch scr; bool fb = false;
            if (mode == 4) m = (k.b_h != 0.0) ? mm::crystal_point_aset<true>(k, a, scr, &it, &fb) : mm::crystal_point_aset<false>(k, a, scr, &it, &fb);
            else if (mode == 3 && k.m_int && k.b_h == 0.0) m = mm::crystal_point_phased(k, a, scr, &it, &fb);
            else if (mode >= 2) m = (k.b_h != 0.0) ? mm::crystal_point_sparse<true>(k, a, scr, &it, &fb) : mm::crystal_point_sparse<false>(k, a, scr, &it, &fb);
            else m = (k.b_h != 0.0) ? mm::crystal_point_fast<true>(k, a, scr, &it, &fb) : mm::crystal_point_fast<false>(k, a, scr, &it, &fb);
            if (fb) { ++nfb; m = mm::crystal_point(k, a, &it); }
        }
        active[i] = (uint16_t)m; iters[i] = (uint16_t)it;
    }
    if (n_fallback) *n_fallback = nfb;
    return 0;
}
int twin_crystal(const MMCrystal* p, const double* deps, double dt, const double* sin, double* sig, double* tan, double* sout, uint16_t* active,
                 uint16_t* iters, double tol, int max_iter, int64_t N) {
    mm::CrystalK k;
    mm::make_crystal_k(k, p->E, p->nu, p->tau_y, p->H_kin, p->alpha_inf, p->t_star, p->sigma_c, p->m, p->MS, p->H, dt, tol, max_iter);
    for (int64_t i = 0; i < N; ++i) {
        HostAcc<2, 3> a{{deps, sin}, {sig, tan, sout}, N, i};
        int it = 0; unsigned m = mm::crystal_point(k, a, &it);
        active[i] = (uint16_t)m; iters[i] = (uint16_t)it;
    }
    return 0;
}

}  // extern "C"
template <int KIND> static void twin_wrapped_plastic_k(const mm::PlasticK& k, const double* eps, const double* sin, double* sig, double* tan, double* sout,
                                                       uint8_t* flags, uint16_t* iters, uint8_t* oiters, double wtol, int wmax, int64_t N) {
    for (int64_t i = 0; i < N; ++i) {
        HostAcc<2, 3> a{{eps, sin}, {sig, tan, sout}, N, i};
        int oit = 0, iit = 0;
        int f = mm::wrapped_point<KIND, 14>([&](mm::RegIO<14>& io, int* it) { return mm::plastic_point(k, io, it); }, a, wtol, wmax, &oit, &iit);
        flags[i] = (uint8_t)f; iters[i] = (uint16_t)iit; oiters[i] = (uint8_t)oit;
    }
}
// the state machine of the lane-persistent kernel (mm_wrap_lp.cuh), one point at a time: sections entered in the kernel's order
struct HostColumn { double* v; double& operator()(int j) const { return v[j]; } };
static uint8_t* g_lp_trace = nullptr;      // optional [N][12]: Newton passes of every inner solve of a point (255 = no such evaluation), for scheduling studies
extern "C" void twin_wrapped_plastic_lp_set_trace(uint8_t* t) { g_lp_trace = t; }
template <int KIND> static void twin_wrapped_plastic_lp_k(const mm::PlasticK& k, const double* eps, const double* sin, double* sig, double* tan, double* sout,
                                                          uint8_t* flags, uint16_t* iters, uint8_t* oiters, double wtol, int wmax, int64_t N) {
    for (int64_t i = 0; i < N; ++i) {
        int ev = -1;
        if (g_lp_trace) memset(g_lp_trace + 12 * i, 255, 12);
        HostAcc<2, 3> a{{eps, sin}, {sig, tan, sout}, N, i};
        double col[mm::J2S_SLOTS];
        HostColumn scr{col};
        mm::J2Lane L;
        mm::j2lp_load<KIND>(a, scr, L);
        while (L.phase != mm::J2P_IDLE) {
            if (L.phase == mm::J2P_TRIAL) { mm::j2lp_trial(k, scr, L); if (g_lp_trace && ++ev < 12) g_lp_trace[12 * i + ev] = 0; }
            if (L.phase == mm::J2P_NEWTON) { mm::j2lp_pass(k, scr, L); if (g_lp_trace && ev < 12) ++g_lp_trace[12 * i + ev]; }
            if (L.phase == mm::J2P_OUTER) mm::j2lp_outer<KIND>(k, scr, L, wtol, wmax);
            if (L.phase == mm::J2P_FINAL) {
                if (!mm::j2lp_final<KIND>(k, scr, L, a)) L.flag |= 0x40;
                flags[i] = (uint8_t)L.flag; iters[i] = (uint16_t)L.iters; oiters[i] = (uint8_t)L.oit;
                L.phase = mm::J2P_IDLE;
            }
        }
    }
}
extern "C" int twin_wrapped_plastic_lp(int kind, const MMPlastic* p, const double* eps, const double* sin, double* sig, double* tan, double* sout, uint8_t* flags,
                         uint16_t* iters, uint8_t* oiters, double tol, int max_iter, double wtol, int wmax, int64_t N) {
    mm::PlasticK k = mm::make_plastic_k(p->E, p->nu, p->sigma_y, p->H, p->r, p->kappa_inf, p->alpha_inf, tol, max_iter);
    switch (kind) {
        case 3: twin_wrapped_plastic_lp_k<3>(k, eps, sin, sig, tan, sout, flags, iters, oiters, wtol, wmax, N); break;
        case 4: twin_wrapped_plastic_lp_k<4>(k, eps, sin, sig, tan, sout, flags, iters, oiters, wtol, wmax, N); break;
        case 5: twin_wrapped_plastic_lp_k<5>(k, eps, sin, sig, tan, sout, flags, iters, oiters, wtol, wmax, N); break;
        default: return -1;
    }
    return 0;
}
extern "C" int twin_wrapped_plastic(int kind, const MMPlastic* p, const double* eps, const double* sin, double* sig, double* tan, double* sout, uint8_t* flags,
                         uint16_t* iters, uint8_t* oiters, double tol, int max_iter, double wtol, int wmax, int64_t N) {
    mm::PlasticK k = mm::make_plastic_k(p->E, p->nu, p->sigma_y, p->H, p->r, p->kappa_inf, p->alpha_inf, tol, max_iter);
    switch (kind) {
        case 1: twin_wrapped_plastic_k<1>(k, eps, sin, sig, tan, sout, flags, iters, oiters, wtol, wmax, N); break;
        case 2: twin_wrapped_plastic_k<2>(k, eps, sin, sig, tan, sout, flags, iters, oiters, wtol, wmax, N); break;
        case 3: twin_wrapped_plastic_k<3>(k, eps, sin, sig, tan, sout, flags, iters, oiters, wtol, wmax, N); break;
        case 4: twin_wrapped_plastic_k<4>(k, eps, sin, sig, tan, sout, flags, iters, oiters, wtol, wmax, N); break;
        case 5: twin_wrapped_plastic_k<5>(k, eps, sin, sig, tan, sout, flags, iters, oiters, wtol, wmax, N); break;
        default: return -1;
    }
    return 0;
}
template <int KIND> static void twin_wrapped_crystal_k(const mm::CrystalK& k, const double* deps, const double* sin, double* sig, double* tan, double* sout,
                                                       uint16_t* active, uint16_t* iters, uint8_t* oiters, double wtol, int wmax, int64_t N) {
    for (int64_t i = 0; i < N; ++i) {
        HostAcc<2, 3> a{{deps, sin}, {sig, tan, sout}, N, i};
        int oit = 0, iit = 0; unsigned mask = 0;
        HostScratch scr;
        int f = mm::wrapped_point<KIND, 42>(
            [&](mm::RegIO<42>& io, int* it) {
                bool fb = !k.h_struct;
                if (!fb) mask = (k.b_h != 0.0) ? mm::crystal_point_fast<true>(k, io, scr, it, &fb) : mm::crystal_point_fast<false>(k, io, scr, it, &fb);
                if (fb) mask = mm::crystal_point(k, io, it);
                return (mask & 0x8000u) ? 0x80 : 0;
            },
            a, wtol, wmax, &oit, &iit);
        if (f & 0x40) mask |= 0x2000u;
        active[i] = (uint16_t)mask; iters[i] = (uint16_t)(iit > 65535 ? 65535 : iit); oiters[i] = (uint8_t)oit;
    }
}
extern "C" int twin_wrapped_crystal(int kind, const MMCrystal* p, const double* deps, double dt, const double* sin, double* sig, double* tan, double* sout,
                         uint16_t* active, uint16_t* iters, uint8_t* oiters, double tol, int max_iter, double wtol, int wmax, int64_t N) {
    mm::CrystalK k;
    mm::make_crystal_k(k, p->E, p->nu, p->tau_y, p->H_kin, p->alpha_inf, p->t_star, p->sigma_c, p->m, p->MS, p->H, dt, tol, max_iter);
    switch (kind) {
        case 1: twin_wrapped_crystal_k<1>(k, deps, sin, sig, tan, sout, active, iters, oiters, wtol, wmax, N); break;
        case 2: twin_wrapped_crystal_k<2>(k, deps, sin, sig, tan, sout, active, iters, oiters, wtol, wmax, N); break;
        case 3: twin_wrapped_crystal_k<3>(k, deps, sin, sig, tan, sout, active, iters, oiters, wtol, wmax, N); break;
        case 4: twin_wrapped_crystal_k<4>(k, deps, sin, sig, tan, sout, active, iters, oiters, wtol, wmax, N); break;
        case 5: twin_wrapped_crystal_k<5>(k, deps, sin, sig, tan, sout, active, iters, oiters, wtol, wmax, N); break;
        default: return -1;
    }
    return 0;
}
template <int KIND> static void twin_wrapped_elastic_k(const mm::ElasticK& k, const double* eps, double* sig, double* tan, uint8_t* oiters, int64_t N) {
    for (int64_t i = 0; i < N; ++i) {
        HostAcc<1, 2> a{{eps}, {sig, tan}, N, i};
        int oit = 0, iit = 0;
        mm::wrapped_point<KIND, 0>([&](mm::RegIO<0>& io, int*) { mm::linear_elastic_point(k, io); return 0; }, a, 1e-8, 10, &oit, &iit);
        oiters[i] = (uint8_t)oit;
    }
}
extern "C" int twin_wrapped_elastic(int kind, const MMLinearElastic* p, const double* eps, double* sig, double* tan, uint8_t* oiters, int64_t N) {
    mm::ElasticK k = mm::make_elastic_k(p->E, p->nu);
    switch (kind) {
        case 1: twin_wrapped_elastic_k<1>(k, eps, sig, tan, oiters, N); break;
        case 2: twin_wrapped_elastic_k<2>(k, eps, sig, tan, oiters, N); break;
        case 3: twin_wrapped_elastic_k<3>(k, eps, sig, tan, oiters, N); break;
        case 4: twin_wrapped_elastic_k<4>(k, eps, sig, tan, oiters, N); break;
        case 5: twin_wrapped_elastic_k<5>(k, eps, sig, tan, oiters, N); break;
        default: return -1;
    }
    return 0;
}

template <int MODEL> static int twin_hyper_ex_m(const mm::HyperK& k, const double* strain, int kind, int tan, double* S, double* T, int64_t N) {
    for (int64_t i = 0; i < N; ++i) {
        HostAcc<1, 2> a{{strain}, {S, T}, N, i};
        if (tan == 2) mm::hyper_point_ex<MODEL, 1, 2>(k, a);
        else if (tan == 1) { if (kind == 1) mm::hyper_point_ex<MODEL, 1, 1>(k, a); else if (kind == 2) mm::hyper_point_ex<MODEL, 2, 1>(k, a); else mm::hyper_point_ex<MODEL, 0, 1>(k, a); }
        else { if (kind == 1) mm::hyper_point_ex<MODEL, 1, 0>(k, a); else if (kind == 2) mm::hyper_point_ex<MODEL, 2, 0>(k, a); else mm::hyper_point_ex<MODEL, 0, 0>(k, a); }
    }
    return 0;
}
extern "C" int twin_hyper_ex(int model, const MMHyper* p, const double* strain, int kind, int tan, double* S, double* T, int64_t N) {
    mm::HyperK k{p->lambda, p->mu, p->c2, p->c3};
    return model == 0 ? twin_hyper_ex_m<0>(k, strain, kind, tan, S, T, N) : (model == 1 ? twin_hyper_ex_m<1>(k, strain, kind, tan, S, T, N) : twin_hyper_ex_m<2>(k, strain, kind, tan, S, T, N));
}

// ---- TransverselyIsotropic and strain energies ----
extern "C" int twin_transviso(const MMTransverselyIsotropic* p, const double* eps, const double* a3, double* sig, double* tan, int64_t N) {
    mm::TransvIsoK k{p->L_perp, p->L_par, p->M_par, p->G_perp, p->G_par};
    for (int64_t i = 0; i < N; ++i) { HostAcc<2, 2> a{{eps, a3}, {sig, tan}, N, i}; mm::transviso_point(k, a); }
    return 0;
}
extern "C" int twin_linear_elastic_energy(const MMLinearElastic* p, const double* eps, double* psi, int64_t N) {
    mm::ElasticK k = mm::make_elastic_k(p->E, p->nu);
    for (int64_t i = 0; i < N; ++i) { HostAcc<1, 1> a{{eps}, {psi}, N, i}; mm::linear_elastic_energy_point(k, a); }
    return 0;
}
template <int MODEL> static void twin_hyper_energy_m(const mm::HyperK& k, const double* strain, int kind, double* psi, int64_t N) {
    for (int64_t i = 0; i < N; ++i) {
        HostAcc<1, 1> a{{strain}, {psi}, N, i};
        if (kind == 1) mm::hyper_energy_point<MODEL, 1>(k, a); else if (kind == 2) mm::hyper_energy_point<MODEL, 2>(k, a); else mm::hyper_energy_point<MODEL, 0>(k, a);
    }
}
extern "C" int twin_hyper_energy(int model, const MMHyper* p, const double* strain, int kind, double* psi, int64_t N) {
    mm::HyperK k{p->lambda, p->mu, p->c2, p->c3};
    if (model == 0) twin_hyper_energy_m<0>(k, strain, kind, psi, N); else if (model == 1) twin_hyper_energy_m<1>(k, strain, kind, psi, N); else twin_hyper_energy_m<2>(k, strain, kind, psi, N);
    return 0;
}

// ---- dimension wrappers around TransverselyIsotropic ----
namespace { template <class Acc> struct NoStateOutH {
    Acc& a;
    template <int A> double in(int c) const { return a.template in<A>(c); }
    template <int A> void out(int c, double v) { if (A < 2) a.template out<(A < 2 ? A : 0)>(c, v); }
    template <int A> bool has_out() const { return A < 2 ? a.template has_out<(A < 2 ? A : 0)>() : false; }
}; }
template <int KIND> static void twin_wrapped_transviso_k(const mm::TransvIsoK& k, const double* eps, const double* a3, double* sig, double* tan, uint8_t* oiters, int64_t N) {
    for (int64_t i = 0; i < N; ++i) {
        HostAcc<2, 2> a{{eps, a3}, {sig, tan}, N, i};
        NoStateOutH<HostAcc<2, 2>> na{a};
        int oit = 0, iit = 0;
        mm::wrapped_point<KIND, 3>([&](mm::RegIO<3>& io, int*) { mm::transviso_point(k, io); return 0; }, na, 1e-8, 10, &oit, &iit);
        oiters[i] = (uint8_t)oit;
    }
}
extern "C" int twin_wrapped_transviso(int kind, const MMTransverselyIsotropic* p, const double* eps, const double* a3, double* sig, double* tan, uint8_t* oiters, int64_t N) {
    mm::TransvIsoK k{p->L_perp, p->L_par, p->M_par, p->G_perp, p->G_par};
    switch (kind) {
        case 1: twin_wrapped_transviso_k<1>(k, eps, a3, sig, tan, oiters, N); break;
        case 2: twin_wrapped_transviso_k<2>(k, eps, a3, sig, tan, oiters, N); break;
        case 3: twin_wrapped_transviso_k<3>(k, eps, a3, sig, tan, oiters, N); break;
        case 4: twin_wrapped_transviso_k<4>(k, eps, a3, sig, tan, oiters, N); break;
        case 5: twin_wrapped_transviso_k<5>(k, eps, a3, sig, tan, oiters, N); break;
        default: return -1;
    }
    return 0;
}

// instrumentation of the sparse crystal pass: how many systems cost a full chain per Newton pass
static long long g_stats_passes = 0, g_stats_nz = 0, g_stats_cand = 0, g_stats_hist[13] = {0};
// optional per-pass trace (tools/crystal_sched_sim.py): one byte per Newton pass = candidate count, 255 = start of a point
static uint8_t* g_trace = nullptr; static long long g_trace_cap = 0, g_trace_len = 0;
extern "C" void twin_crystal_trace(uint8_t* buf, long long cap) { g_trace = buf; g_trace_cap = cap; g_trace_len = 0; }
extern "C" long long twin_crystal_trace_len() { return g_trace_len; }
static inline void trace_put(uint8_t v) { if (g_trace && g_trace_len < g_trace_cap) g_trace[g_trace_len++] = v; }
namespace mm { void mm_crystal_stats_pass(unsigned nz, unsigned cand) { trace_put((uint8_t)__builtin_popcount(cand)); ++g_stats_passes; g_stats_nz += __builtin_popcount(nz); const int c = __builtin_popcount(cand); g_stats_cand += c; ++g_stats_hist[c]; } }
extern "C" void twin_crystal_stats(long long* out16, int reset) {
    out16[0] = g_stats_passes; out16[1] = g_stats_nz; out16[2] = g_stats_cand;
    for (int i = 0; i < 13; ++i) out16[3 + i] = g_stats_hist[i];
    if (reset) { g_stats_passes = g_stats_nz = g_stats_cand = 0; for (int i = 0; i < 13; ++i) g_stats_hist[i] = 0; }
}
