// ORACLE — TEST INFRASTRUCTURE ONLY (see mmo_tensor.hpp header).
//
// Batched CPU entry points (`mmo_*`) with the same argument meaning as the product's C ABI
// (include/mmb200.h, of which only the POD parameter structs are used here), host pointers,
// OpenMP over points.  This is the checker for the CUDA path and the CPU baseline that bench.py
// reports; the product never links or loads it.
#include <cstdint>
#include <vector>
#include <cstring>
#include <cmath>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/mmb200.h"
#include "mmo_models.hpp"

using namespace mmo;

namespace {
struct View {  // component accessor for either layout
    double* p; int64_t N; int nc; int layout;
    inline double& at(int64_t i, int c) const { return layout == MM_LAYOUT_SOA ? p[(int64_t)c * N + i] : p[i * nc + c]; }
};
inline View V(const double* p, int64_t N, int nc, int layout) { return View{const_cast<double*>(p), N, nc, layout}; }

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
}  // namespace

extern "C" {

int mmo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void mmo_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int mmo_elastic_tangent_3d(double E, double nu, double* out36) {
    Sym4<double> C = elastic_tangent_3D(E, nu);
    std::memcpy(out36, C.a, sizeof(C.a));
    return 0;
}

int mmo_slipsystems(int structure, const double rodrigues[3], double* planes, double* dirs) {
    double R[9]; rodrigues_matrix(rodrigues, R);
    double pl[12][3], di[12][3];
    int n = slipsystems(structure == MM_FCC ? MMO_FCC : MMO_BCC, R, pl, di);
    for (int s = 0; s < n; ++s) for (int c = 0; c < 3; ++c) { planes[s * 3 + c] = pl[s][c]; dirs[s * 3 + c] = di[s][c]; }
    return n;
}

int mmo_crystal_setup(MMCrystal* p, const double* planes, const double* dirs) {
    double pl[12][3], di[12][3];
    for (int s = 0; s < 12; ++s) for (int c = 0; c < 3; ++c) { pl[s][c] = planes[s * 3 + c]; di[s][c] = dirs[s * 3 + c]; }
    CrystalM<12> c = make_crystal<12>(p->E, p->nu, p->tau_y, p->H_iso, p->H_kin, p->q, p->alpha_inf, p->t_star, p->sigma_c, p->m, pl, di);
    p->nslip = 12;
    for (int s = 0; s < 12; ++s) for (int k = 0; k < 9; ++k) p->MS[s][k] = c.MS[s].a[k];
    for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) p->H[i][j] = c.H[i + 12 * j];
    return 0;
}

int mmo_xu_needleman_setup(double sigma_max, double tau_max, double Phi_n, double Phi_t, double Delta_n_star, MMXuNeedleman* out) {
    XuM m = make_xu(sigma_max, tau_max, Phi_n, Phi_t, Delta_n_star);
    out->Phi_n = m.Phi_n; out->Phi_t = m.Phi_t; out->delta_n = m.delta_n; out->delta_t = m.delta_t; out->Delta_n_star = m.Delta_n_star;
    return 0;
}

int mmo_linear_elastic(const MMLinearElastic* p, const double* eps, double* sig, double* tan_or_null, int64_t N, int layout) {
    LinearElasticM m = make_linear_elastic(p->E, p->nu);
    View ve = V(eps, N, 6, layout), vs = V(sig, N, 6, layout), vt = V(tan_or_null, N, 36, layout);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        Sym2<double> e, s; Sym4<double> t;
        for (int c = 0; c < 6; ++c) e.a[c] = ve.at(i, c);
        material_response(m, e, s, t);
        for (int c = 0; c < 6; ++c) vs.at(i, c) = s.a[c];
        if (tan_or_null) for (int c = 0; c < 36; ++c) vt.at(i, c) = t.a[c];
    }
    return 0;
}

int mmo_plastic(const MMPlastic* p, const double* eps, const double* state_in, double* sig, double* tan, double* state_out,
                uint8_t* flags, uint16_t* iters, int32_t* n_fail, const MMNewtonOpts* opts, int64_t N, int layout) {
    PlasticM m = make_plastic(p->E, p->nu, p->sigma_y, p->H, p->r, p->kappa_inf, p->alpha_inf);
    NewtonOpts o{opts ? opts->tol : 1e-8, opts ? opts->max_iter : 1000};
    View ve = V(eps, N, 6, layout), vi = V(state_in, N, 14, layout), vs = V(sig, N, 6, layout), vt = V(tan, N, 36, layout),
         vo = V(state_out, N, 14, layout);
    int64_t fails = 0;
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : fails)
    for (int64_t i = 0; i < N; ++i) {
        Sym2<double> e, s; Sym4<double> t; PlasticState st, sn;
        for (int c = 0; c < 6; ++c) e.a[c] = ve.at(i, c);
        for (int c = 0; c < 6; ++c) st.ep.a[c] = vi.at(i, c);
        st.kappa = vi.at(i, 6);
        for (int c = 0; c < 6; ++c) st.alpha.a[c] = vi.at(i, 7 + c);
        st.mu = vi.at(i, 13);
        int flag = 0, it = 0;
        int status = material_response(m, e, st, o, s, t, sn, flag, it);
        if (status) { ++fails; flag |= MM_FLAG_FAILED; }
        for (int c = 0; c < 6; ++c) vs.at(i, c) = s.a[c];
        if (tan) for (int c = 0; c < 36; ++c) vt.at(i, c) = t.a[c];
        for (int c = 0; c < 6; ++c) vo.at(i, c) = sn.ep.a[c];
        vo.at(i, 6) = sn.kappa;
        for (int c = 0; c < 6; ++c) vo.at(i, 7 + c) = sn.alpha.a[c];
        vo.at(i, 13) = sn.mu;
        if (flags) flags[i] = (uint8_t)flag;
        if (iters) iters[i] = (uint16_t)(it > 65535 ? 65535 : it);
    }
    if (n_fail) *n_fail += (int32_t)fails;
    return 0;
}

int mmo_hyper(int model, const MMHyper* p, const double* strain, int strain_kind, double* S, double* dSdC, int64_t N, int layout) {
    HyperM m{p->lambda, p->mu, p->c2, p->c3};
    int nin = strain_kind == MM_STRAIN_F ? 9 : 6;
    View vi = V(strain, N, nin, layout), vs = V(S, N, 6, layout), vt = V(dSdC, N, 36, layout);
    int om = model == MM_NEOHOOK ? MMO_NEOHOOK : (model == MM_YEOH ? MMO_YEOH : MMO_STVENANT);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        Sym2<double> C, s; Sym4<double> t;
        if (strain_kind == MM_STRAIN_F) { Ten2<double> F; for (int c = 0; c < 9; ++c) F.a[c] = vi.at(i, c); C = tdot(F); }
        else for (int c = 0; c < 6; ++c) C.a[c] = vi.at(i, c);
        hyper_response(om, m, C, s, t);
        for (int c = 0; c < 6; ++c) vs.at(i, c) = s.a[c];
        if (dSdC) for (int c = 0; c < 36; ++c) vt.at(i, c) = t.a[c];
    }
    return 0;
}

// AD cross-check used by the oracle's own tests (mirrors test_neohook.jl:13-16): from C, single point
int mmo_hyper_ad(int model, const MMHyper* p, const double* C6, double* psi, double* S6, double* dSdC36) {
    HyperM m{p->lambda, p->mu, p->c2, p->c3};
    Sym2<double> C, s; Sym4<double> t;
    for (int c = 0; c < 6; ++c) C.a[c] = C6[c];
    int om = model == MM_NEOHOOK ? MMO_NEOHOOK : (model == MM_YEOH ? MMO_YEOH : MMO_STVENANT);
    hyper_response_ad(om, m, C, *psi, s, t);
    std::memcpy(S6, s.a, sizeof(s.a)); std::memcpy(dSdC36, t.a, sizeof(t.a));
    return 0;
}

int mmo_crystal(const MMCrystal* p, const double* deps, double dt, const double* state_in, double* sig, double* tan, double* state_out,
                uint16_t* active, uint16_t* iters, int32_t* n_fail, const MMNewtonOpts* opts, int64_t N, int layout) {
    if (p->nslip != 12) return MM_ERR_ARG;
    CrystalM<12> m;
    m.E = p->E; m.nu = p->nu; m.tau_y = p->tau_y; m.H_iso = p->H_iso; m.H_kin = p->H_kin; m.q = p->q; m.alpha_inf = p->alpha_inf;
    m.t_star = p->t_star; m.sigma_c = p->sigma_c; m.m = p->m;
    m.Ee = elastic_tangent_3D(p->E, p->nu);
    for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) m.H[i + 12 * j] = p->H[i][j];
    for (int s = 0; s < 12; ++s) for (int k = 0; k < 9; ++k) m.MS[s].a[k] = p->MS[s][k];
    double tol = opts ? opts->tol : 1e-8; int max_iter = opts ? opts->max_iter : 100;
    View ve = V(deps, N, 6, layout), vi = V(state_in, N, 42, layout), vs = V(sig, N, 6, layout), vt = V(tan, N, 36, layout),
         vo = V(state_out, N, 42, layout);
    int64_t fails = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : fails)
    for (int64_t i = 0; i < N; ++i) {
        Sym2<double> e, s; Sym4<double> t; CrystalState<12> st, sn;
        for (int c = 0; c < 6; ++c) e.a[c] = ve.at(i, c);
        for (int c = 0; c < 6; ++c) st.sigma.a[c] = vi.at(i, c);
        for (int k = 0; k < 12; ++k) { st.kappa[k] = vi.at(i, 6 + k); st.alpha[k] = vi.at(i, 18 + k); st.mu[k] = vi.at(i, 30 + k); }
        unsigned act = 0; int it = 0;
        int status = material_response<12>(m, e, st, dt, tol, max_iter, s, t, sn, act, it);
        if (status) { ++fails; act |= MM_ACTIVE_FAILED; }
        for (int c = 0; c < 6; ++c) vs.at(i, c) = s.a[c];
        if (tan) for (int c = 0; c < 36; ++c) vt.at(i, c) = t.a[c];
        for (int c = 0; c < 6; ++c) vo.at(i, c) = sn.sigma.a[c];
        for (int k = 0; k < 12; ++k) { vo.at(i, 6 + k) = sn.kappa[k]; vo.at(i, 18 + k) = sn.alpha[k]; vo.at(i, 30 + k) = sn.mu[k]; }
        if (active) active[i] = (uint16_t)act;
        if (iters) iters[i] = (uint16_t)(it > 65535 ? 65535 : it);
    }
    if (n_fail) *n_fail += (int32_t)fails;
    return 0;
}

int mmo_xu_needleman(const MMXuNeedleman* p, const double* jump, double* T, double* dTdD, int dim, int64_t N, int layout) {
    if (dim != 2 && dim != 3) return MM_ERR_ARG;
    XuM m{p->Phi_n, p->Phi_t, p->delta_n, p->delta_t, p->Delta_n_star};
    View vj = V(jump, N, dim, layout), vT = V(T, N, dim, layout), vH = V(dTdD, N, dim * dim, layout);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        double d[3], t[3], h[9];
        for (int c = 0; c < dim; ++c) d[c] = vj.at(i, c);
        if (dim == 3) xu_response<3>(m, d, t, h); else xu_response<2>(m, d, t, h);
        for (int c = 0; c < dim; ++c) vT.at(i, c) = t[c];
        if (dTdD) for (int c = 0; c < dim * dim; ++c) vH.at(i, c) = h[c];
    }
    return 0;
}

// solve KATs (test_newton_solver.jl): which = 0: f(x) = 2x + 1, which = 1: f(x) = exp(x) - 1
int mmo_solve_scalar_kat(int which, double x0, double* x, double* dfdx, int* loop_index) {
    bool ok;
    if (which == 0) ok = solve_scalar([](Dual<1> v) { return 2.0 * v + 1.0; }, x0, 1e-8, 100, *x, *dfdx, *loop_index);
    else            ok = solve_scalar([](Dual<1> v) { return exp_(v) - 1.0; }, x0, 1e-8, 100, *x, *dfdx, *loop_index);
    return ok ? 0 : 1;
}

// energies for the KATs (test_neohook.jl:5-6, test_yeoh.jl:6-7)
double mmo_strain_energy(int model, const MMHyper* p, const double* C6) {
    HyperM m{p->lambda, p->mu, p->c2, p->c3};
    Sym2<double> C; for (int c = 0; c < 6; ++c) C.a[c] = C6[c];
    int om = model == MM_NEOHOOK ? MMO_NEOHOOK : (model == MM_YEOH ? MMO_YEOH : MMO_STVENANT);
    return strain_energy<double>(om, m, C);
}

int mmo_synth_uniform(double* out, int64_t N, int nc, int layout, uint64_t seed, const double* lo, const double* hi) {
    View v = V(out, N, nc, layout);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i)
        for (int c = 0; c < nc; ++c) {
            double u = (double)(splitmix64(seed + (uint64_t)i * (uint64_t)nc + (uint64_t)c) >> 11) * (1.0 / 9007199254740992.0);
            v.at(i, c) = lo[c] + (hi[c] - lo[c]) * u;
        }
    return 0;
}

// ---- dimension wrappers (reference src/wrappers.jl) ------------------------------------------------------------
int mmo_wrapped_linear_elastic(int kind, const MMLinearElastic* p, const double* eps_red, double* sig_red, double* tan_red,
                               uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* wopts, int64_t N, int layout) {
    LinearElasticM m = make_linear_elastic(p->E, p->nu);
    const int nc = wrap_ncomp(kind);
    double tol = wopts ? wopts->tol : 1e-8; int max_iter = wopts ? wopts->max_iter : 10;
    View ve = V(eps_red, N, nc, layout), vs = V(sig_red, N, nc, layout), vt = V(tan_red, N, nc * nc, layout);
    int64_t fails = 0;
#pragma omp parallel for schedule(static) reduction(+ : fails)
    for (int64_t i = 0; i < N; ++i) {
        double e[6], s[6], t[36]; int oit = 0;
        for (int c = 0; c < nc; ++c) e[c] = ve.at(i, c);
        auto inner = [&](const Sym2<double>& eps3, Sym2<double>& sig, Sym4<double>& tan) { material_response(m, eps3, sig, tan); return 0; };
        int st = wrapped_response(kind, inner, e, tol, max_iter, s, t, oit);
        if (st) ++fails;
        for (int c = 0; c < nc; ++c) vs.at(i, c) = s[c];
        if (tan_red) for (int c = 0; c < nc * nc; ++c) vt.at(i, c) = t[c];
        if (outer_iters) outer_iters[i] = (uint8_t)oit;
    }
    if (n_fail) *n_fail += (int32_t)fails;
    return 0;
}

int mmo_wrapped_plastic(int kind, const MMPlastic* p, const double* eps_red, const double* state_in, double* sig_red, double* tan_red,
                        double* state_out, uint8_t* flags, uint16_t* iters, uint8_t* outer_iters, int32_t* n_fail,
                        const MMNewtonOpts* opts, const MMNewtonOpts* wopts, int64_t N, int layout) {
    PlasticM m = make_plastic(p->E, p->nu, p->sigma_y, p->H, p->r, p->kappa_inf, p->alpha_inf);
    NewtonOpts o{opts ? opts->tol : 1e-8, opts ? opts->max_iter : 1000};
    const int nc = wrap_ncomp(kind);
    double tol = wopts ? wopts->tol : 1e-8; int max_iter = wopts ? wopts->max_iter : 10;
    View ve = V(eps_red, N, nc, layout), vi = V(state_in, N, 14, layout), vs = V(sig_red, N, nc, layout), vt = V(tan_red, N, nc * nc, layout),
         vo = V(state_out, N, 14, layout);
    int64_t fails = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : fails)
    for (int64_t i = 0; i < N; ++i) {
        double e[6], s[6], t[36]; int oit = 0;
        for (int c = 0; c < 6; ++c) { s[c] = 0; }
        for (int c = 0; c < 36; ++c) t[c] = 0;
        for (int c = 0; c < nc; ++c) e[c] = ve.at(i, c);
        PlasticState st, sn;
        for (int c = 0; c < 6; ++c) st.ep.a[c] = vi.at(i, c);
        st.kappa = vi.at(i, 6);
        for (int c = 0; c < 6; ++c) st.alpha.a[c] = vi.at(i, 7 + c);
        st.mu = vi.at(i, 13);
        sn = st;
        int flag = 0, it = 0;
        auto inner = [&](const Sym2<double>& eps3, Sym2<double>& sig, Sym4<double>& tan) { return material_response(m, eps3, st, o, sig, tan, sn, flag, it); };
        int status = wrapped_response(kind, inner, e, tol, max_iter, s, t, oit);
        if (status == 1) flag |= MM_FLAG_FAILED;
        if (status == 2) flag |= 0x40;
        if (status) ++fails;
        for (int c = 0; c < nc; ++c) vs.at(i, c) = s[c];
        if (tan_red) for (int c = 0; c < nc * nc; ++c) vt.at(i, c) = t[c];
        for (int c = 0; c < 6; ++c) vo.at(i, c) = sn.ep.a[c];
        vo.at(i, 6) = sn.kappa;
        for (int c = 0; c < 6; ++c) vo.at(i, 7 + c) = sn.alpha.a[c];
        vo.at(i, 13) = sn.mu;
        if (flags) flags[i] = (uint8_t)flag;
        if (iters) iters[i] = (uint16_t)(it > 65535 ? 65535 : it);
        if (outer_iters) outer_iters[i] = (uint8_t)oit;
    }
    if (n_fail) *n_fail += (int32_t)fails;
    return 0;
}

int mmo_wrapped_transviso(int kind, const MMTransverselyIsotropic* p, const double* eps_red, const double* a3, double* sig_red, double* tan_red,
                          uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* wopts, int64_t N, int layout) {
    TransvIsoM m{p->L_perp, p->L_par, p->M_par, p->G_perp, p->G_par};
    const int nc = wrap_ncomp(kind);
    double tol = wopts ? wopts->tol : 1e-8; int max_iter = wopts ? wopts->max_iter : 10;
    View ve = V(eps_red, N, nc, layout), va = V(a3, N, 3, layout), vs = V(sig_red, N, nc, layout), vt = V(tan_red, N, nc * nc, layout);
    int64_t fails = 0;
#pragma omp parallel for schedule(static) reduction(+ : fails)
    for (int64_t i = 0; i < N; ++i) {
        double e[6], s[6], t[36], a[3]; int oit = 0;
        for (int c = 0; c < nc; ++c) e[c] = ve.at(i, c);
        for (int c = 0; c < 3; ++c) a[c] = va.at(i, c);
        auto inner = [&](const Sym2<double>& eps3, Sym2<double>& sig, Sym4<double>& tan) { material_response(m, eps3, a, sig, tan); return 0; };
        int st = wrapped_response(kind, inner, e, tol, max_iter, s, t, oit);
        if (st) ++fails;
        for (int c = 0; c < nc; ++c) vs.at(i, c) = s[c];
        if (tan_red) for (int c = 0; c < nc * nc; ++c) vt.at(i, c) = t[c];
        if (outer_iters) outer_iters[i] = (uint8_t)oit;
    }
    if (n_fail) *n_fail += (int32_t)fails;
    return 0;
}

int mmo_wrapped_crystal(int kind, const MMCrystal* p, const double* deps_red, double dt, const double* state_in, double* sig_red, double* tan_red,
                        double* state_out, uint16_t* active, uint16_t* iters, uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* opts,
                        const MMNewtonOpts* wopts, int64_t N, int layout) {
    if (p->nslip != 12) return MM_ERR_ARG;
    CrystalM<12> m;
    m.E = p->E; m.nu = p->nu; m.tau_y = p->tau_y; m.H_iso = p->H_iso; m.H_kin = p->H_kin; m.q = p->q; m.alpha_inf = p->alpha_inf;
    m.t_star = p->t_star; m.sigma_c = p->sigma_c; m.m = p->m;
    m.Ee = elastic_tangent_3D(p->E, p->nu);
    for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) m.H[i + 12 * j] = p->H[i][j];
    for (int s = 0; s < 12; ++s) for (int k = 0; k < 9; ++k) m.MS[s].a[k] = p->MS[s][k];
    const double tol = opts ? opts->tol : 1e-8; const int max_iter = opts ? opts->max_iter : 100;
    const double wtol = wopts ? wopts->tol : 1e-8; const int wmax = wopts ? wopts->max_iter : 10;
    const int nc = wrap_ncomp(kind);
    View ve = V(deps_red, N, nc, layout), vi = V(state_in, N, 42, layout), vs = V(sig_red, N, nc, layout), vt = V(tan_red, N, nc * nc, layout),
         vo = V(state_out, N, 42, layout);
    int64_t fails = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : fails)
    for (int64_t i = 0; i < N; ++i) {
        double e[6], s[6] = {0, 0, 0, 0, 0, 0}, t[36]; int oit = 0;
        for (int c = 0; c < 36; ++c) t[c] = 0;
        for (int c = 0; c < nc; ++c) e[c] = ve.at(i, c);
        CrystalState<12> st, sn;
        for (int c = 0; c < 6; ++c) st.sigma.a[c] = vi.at(i, c);
        for (int k = 0; k < 12; ++k) { st.kappa[k] = vi.at(i, 6 + k); st.alpha[k] = vi.at(i, 18 + k); st.mu[k] = vi.at(i, 30 + k); }
        sn = st;
        unsigned act = 0; int it = 0;
        auto inner = [&](const Sym2<double>& eps3, Sym2<double>& sig, Sym4<double>& tan) { return material_response<12>(m, eps3, st, dt, tol, max_iter, sig, tan, sn, act, it); };
        int status = wrapped_response(kind, inner, e, wtol, wmax, s, t, oit);
        if (status == 1) act |= MM_ACTIVE_FAILED;
        if (status == 2) act |= 0x2000u;
        if (status) ++fails;
        for (int c = 0; c < nc; ++c) vs.at(i, c) = s[c];
        if (tan_red) for (int c = 0; c < nc * nc; ++c) vt.at(i, c) = t[c];
        for (int c = 0; c < 6; ++c) vo.at(i, c) = sn.sigma.a[c];
        for (int k = 0; k < 12; ++k) { vo.at(i, 6 + k) = sn.kappa[k]; vo.at(i, 18 + k) = sn.alpha[k]; vo.at(i, 30 + k) = sn.mu[k]; }
        if (active) active[i] = (uint16_t)act;
        if (iters) iters[i] = (uint16_t)(it > 65535 ? 65535 : it);
        if (outer_iters) outer_iters[i] = (uint8_t)oit;
    }
    if (n_fail) *n_fail += (int32_t)fails;
    return 0;
}

// ---- the caller's quadrature loop around material_response (Ferrite.jl-style assemble_cell!) --------
//   ε = symmetric(Σ_a u_a ⊗ ∇N_a);  σ, D, state_new = material_response(m, ε, state);  fe[a] += (σ ⋅ ∇N_a) dΩ
// model: 0 LinearElastic (params = MMLinearElastic), 1 Plastic (params = MMPlastic), 2 TransverselyIsotropic (params =
// MMTransverselyIsotropic, state_in = a3[3][P]).  Arrays as include/mmb200.h mm_cells_*.
int mmo_cells(int model, const void* params, int nn, int nq, const double* dNdx, const double* dOmega, const double* ue, const double* state_in,
              double* fe, double* ke, double* sig_out, double* tan_out, double* state_out, uint8_t* flags, uint16_t* iters, int32_t* n_fail, const MMNewtonOpts* opts,
              int64_t ncell) {
    const int64_t P = ncell * nq;
    LinearElasticM le{}; PlasticM pm{}; TransvIsoM tm{};
    if (model == 0) { const MMLinearElastic* p = (const MMLinearElastic*)params; le = make_linear_elastic(p->E, p->nu); }
    else if (model == 2) { const MMTransverselyIsotropic* p = (const MMTransverselyIsotropic*)params; tm = TransvIsoM{p->L_perp, p->L_par, p->M_par, p->G_perp, p->G_par}; }
    else { const MMPlastic* p = (const MMPlastic*)params; pm = make_plastic(p->E, p->nu, p->sigma_y, p->H, p->r, p->kappa_inf, p->alpha_inf); }
    NewtonOpts o{opts ? opts->tol : 1e-8, opts ? opts->max_iter : 1000};
    int64_t fails = 0;
#pragma omp parallel for schedule(static) reduction(+ : fails)
    for (int64_t cell = 0; cell < ncell; ++cell) {
        std::vector<double> f(3 * nn, 0.0), K(ke ? 9 * nn * nn : 0, 0.0);
        for (int q = 0; q < nq; ++q) {
            const int64_t i = cell * nq + q;
            Ten2<double> G;                                                  // ∇u = Σ_a u_a ⊗ ∇N_a
            for (int a = 0; a < nn; ++a)
                for (int k = 0; k < 3; ++k)
                    for (int l = 0; l < 3; ++l) G(k, l) = G(k, l) + ue[(int64_t)(3 * a + k) * ncell + cell] * dNdx[(int64_t)(3 * a + l) * P + i];
            Sym2<double> eps, sig; Sym4<double> tan;
            for (int l = 0; l < 3; ++l) for (int k = l; k < 3; ++k) eps(k, l) = 0.5 * (G(k, l) + G(l, k));
            if (model == 0) material_response(le, eps, sig, tan);
            else if (model == 2) {
                double a3[3];
                for (int c = 0; c < 3; ++c) a3[c] = state_in[(int64_t)c * P + i];
                material_response(tm, eps, a3, sig, tan);
            } else {
                PlasticState st, sn;
                for (int c = 0; c < 6; ++c) st.ep.a[c] = state_in[(int64_t)c * P + i];
                st.kappa = state_in[(int64_t)6 * P + i];
                for (int c = 0; c < 6; ++c) st.alpha.a[c] = state_in[(int64_t)(7 + c) * P + i];
                st.mu = state_in[(int64_t)13 * P + i];
                int flag = 0, it = 0;
                int status = material_response(pm, eps, st, o, sig, tan, sn, flag, it);
                if (status) { flag |= MM_FLAG_FAILED; ++fails; }
                for (int c = 0; c < 6; ++c) state_out[(int64_t)c * P + i] = sn.ep.a[c];
                state_out[(int64_t)6 * P + i] = sn.kappa;
                for (int c = 0; c < 6; ++c) state_out[(int64_t)(7 + c) * P + i] = sn.alpha.a[c];
                state_out[(int64_t)13 * P + i] = sn.mu;
                if (flags) flags[i] = (uint8_t)flag;
                if (iters) iters[i] = (uint16_t)(it > 65535 ? 65535 : it);
            }
            if (sig_out) for (int c = 0; c < 6; ++c) sig_out[(int64_t)c * P + i] = sig.a[c];
            if (tan_out) for (int c = 0; c < 36; ++c) tan_out[(int64_t)c * P + i] = tan.a[c];
            for (int a = 0; a < nn; ++a)
                for (int k = 0; k < 3; ++k) {
                    double v = 0;
                    for (int l = 0; l < 3; ++l) v += sig(k, l) * dNdx[(int64_t)(3 * a + l) * P + i];
                    f[3 * a + k] += v * dOmega[i];
                }
            if (ke)                                                          // ke[i,j] += δε_i ⊡ D ⊡ Δε_j dΩ, δε = symmetric(e_k ⊗ ∇N_a)
                for (int a = 0; a < nn; ++a) for (int k = 0; k < 3; ++k) {
                    Sym2<double> de;
                    for (int l = 0; l < 3; ++l) { de(k, l) = 0; }
                    for (int l = 0; l < 3; ++l) de(k, l) = de(k, l) + 0.5 * dNdx[(int64_t)(3 * a + l) * P + i] * (k == l ? 2.0 : 1.0);
                    for (int b = 0; b < nn; ++b) for (int m2 = 0; m2 < 3; ++m2) {
                        Sym2<double> De;
                        for (int l = 0; l < 3; ++l) De(m2, l) = De(m2, l) + 0.5 * dNdx[(int64_t)(3 * b + l) * P + i] * (m2 == l ? 2.0 : 1.0);
                        K[(3 * a + k) * 3 * nn + 3 * b + m2] += dcontract(de, dcontract(tan, De)) * dOmega[i];
                    }
                }
        }
        for (int c = 0; c < 3 * nn; ++c) fe[(int64_t)c * ncell + cell] = f[c];
        if (ke) for (int c = 0; c < 9 * nn * nn; ++c) ke[(int64_t)c * ncell + cell] = K[c];
    }
    if (n_fail) *n_fail += (int32_t)fails;
    return 0;
}

// total-Lagrangian cell loop: F = I + Σ_a u_a ⊗ ∇N_a; (Pᵀ, ∂Pᵀ∂F) = material_response(∂Pᵀ∂F(), m, DeformationGradient(F)) (traits.jl:67-107);
// fe[a][i] += (P ⋅ ∇N_a)_i dΩ with P = (Pᵀ)ᵀ
int mmo_cells_hyper(int model, const MMHyper* p, int nn, int nq, const double* dNdx, const double* dOmega, const double* ue, double* fe, double* Pt_out,
                    double* dP_out, int64_t ncell) {
    const int64_t P = ncell * nq;
    HyperM m{p->lambda, p->mu, p->c2, p->c3};
    int om = model == MM_NEOHOOK ? MMO_NEOHOOK : (model == MM_YEOH ? MMO_YEOH : MMO_STVENANT);
#pragma omp parallel for schedule(static)
    for (int64_t cell = 0; cell < ncell; ++cell) {
        std::vector<double> f(3 * nn, 0.0);
        for (int q = 0; q < nq; ++q) {
            const int64_t i = cell * nq + q;
            double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Pt[9], dP[81];
            for (int a = 0; a < nn; ++a)
                for (int k = 0; k < 3; ++k)
                    for (int l = 0; l < 3; ++l) F[k + 3 * l] += ue[(int64_t)(3 * a + k) * ncell + cell] * dNdx[(int64_t)(3 * a + l) * P + i];
            hyper_response_traits(om, m, F, MMO_STRAIN_F, MMO_TAN_DPDF, Pt, dP);
            if (Pt_out) for (int c = 0; c < 9; ++c) Pt_out[(int64_t)c * P + i] = Pt[c];
            if (dP_out) for (int c = 0; c < 81; ++c) dP_out[(int64_t)c * P + i] = dP[c];
            for (int a = 0; a < nn; ++a)
                for (int k = 0; k < 3; ++k) {
                    double v = 0;
                    for (int l = 0; l < 3; ++l) v += Pt[l + 3 * k] * dNdx[(int64_t)(3 * a + l) * P + i];     // P_kl = Pᵀ_lk
                    f[3 * a + k] += v * dOmega[i];
                }
        }
        for (int c = 0; c < 3 * nn; ++c) fe[(int64_t)c * ncell + cell] = f[c];
    }
    return 0;
}

// ---- trait layer (reference src/traits.jl:34-107): strain kinds C / F / E, tangent kinds ∂S∂C / ∂S∂E / ∂Pᵀ∂F ----------
int mmo_hyper_ex(int model, const MMHyper* p, const double* strain, int strain_kind, int tangent_kind, double* stress, double* tangent,
                 int64_t N, int layout) {
    HyperM m{p->lambda, p->mu, p->c2, p->c3};
    if (tangent_kind == MMO_TAN_DPDF && strain_kind != MMO_STRAIN_F) return MM_ERR_ARG;
    const int nin = strain_kind == MMO_STRAIN_F ? 9 : 6, ns = tangent_kind == MMO_TAN_DPDF ? 9 : 6, nt = tangent_kind == MMO_TAN_DPDF ? 81 : 36;
    View vi = V(strain, N, nin, layout), vs = V(stress, N, ns, layout), vt = V(tangent, N, nt, layout);
    int om = model == MM_NEOHOOK ? MMO_NEOHOOK : (model == MM_YEOH ? MMO_YEOH : MMO_STVENANT);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        double in[9], s[9], t[81];
        for (int c = 0; c < nin; ++c) in[c] = vi.at(i, c);
        hyper_response_traits(om, m, in, strain_kind, tangent_kind, s, t);
        for (int c = 0; c < ns; ++c) vs.at(i, c) = s[c];
        if (tangent) for (int c = 0; c < nt; ++c) vt.at(i, c) = t[c];
    }
    return 0;
}
// AD check (NeoHook): Pᵀ and ∂Pᵀ∂F by dual numbers through F -> S(FᵀF)⋅Fᵀ, single point
int mmo_neohook_dPdF_ad(const MMHyper* p, const double* F9, double* Pt9, double* dP81) {
    HyperM m{p->lambda, p->mu, p->c2, p->c3};
    Ten2<double> F, Pt; Ten4<double> dP;
    for (int c = 0; c < 9; ++c) F.a[c] = F9[c];
    neohook_dPdF_ad(m, F, Pt, dP);
    std::memcpy(Pt9, Pt.a, sizeof(Pt.a)); std::memcpy(dP81, dP.a, sizeof(dP.a));
    return 0;
}

// ---- TransverselyIsotropic + strain energies --------------------------------------
int mmo_transviso_setup(double E_L, double E_T, double G_LT, double nu_LT, double nu_TT, MMTransverselyIsotropic* out) {
    TransvIsoM m = make_transviso(E_L, E_T, G_LT, nu_LT, nu_TT);
    out->L_perp = m.L_perp; out->L_par = m.L_par; out->M_par = m.M_par; out->G_perp = m.G_perp; out->G_par = m.G_par;
    return 0;
}
int mmo_transversely_isotropic(const MMTransverselyIsotropic* p, const double* eps, const double* a3, double* sig, double* tan, int64_t N, int layout) {
    TransvIsoM m{p->L_perp, p->L_par, p->M_par, p->G_perp, p->G_par};
    View ve = V(eps, N, 6, layout), va = V(a3, N, 3, layout), vs = V(sig, N, 6, layout), vt = V(tan, N, 36, layout);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        Sym2<double> e, s; Sym4<double> t; double a[3];
        for (int c = 0; c < 6; ++c) e.a[c] = ve.at(i, c);
        for (int c = 0; c < 3; ++c) a[c] = va.at(i, c);
        material_response(m, e, a, s, t);
        for (int c = 0; c < 6; ++c) vs.at(i, c) = s.a[c];
        if (tan) for (int c = 0; c < 36; ++c) vt.at(i, c) = t.a[c];
    }
    return 0;
}
// elastic_strain_energy_density: model = -1 LinearElastic (p = MMLinearElastic, strain = ε), else hyper model with strain kind C/F/E
int mmo_strain_energy_batch(int model, const void* params, const double* strain, int strain_kind, double* psi, int64_t N, int layout) {
    const int nin = (model >= 0 && strain_kind == MMO_STRAIN_F) ? 9 : 6;
    View vi = V(strain, N, nin, layout);
    if (model < 0) {
        const MMLinearElastic* p = (const MMLinearElastic*)params;
        LinearElasticM m = make_linear_elastic(p->E, p->nu);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) { Sym2<double> e; for (int c = 0; c < 6; ++c) e.a[c] = vi.at(i, c); psi[i] = linear_elastic_energy(m, e); }
        return 0;
    }
    const MMHyper* p = (const MMHyper*)params;
    HyperM m{p->lambda, p->mu, p->c2, p->c3};
    int om = model == MM_NEOHOOK ? MMO_NEOHOOK : (model == MM_YEOH ? MMO_YEOH : MMO_STVENANT);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        Sym2<double> C;
        if (strain_kind == MMO_STRAIN_F) { Ten2<double> F; for (int c = 0; c < 9; ++c) F.a[c] = vi.at(i, c); C = tdot(F); }
        else if (strain_kind == MMO_STRAIN_E) { for (int c = 0; c < 6; ++c) C.a[c] = 2 * vi.at(i, c); C(0, 0) += 1; C(1, 1) += 1; C(2, 2) += 1; }
        else for (int c = 0; c < 6; ++c) C.a[c] = vi.at(i, c);
        psi[i] = strain_energy<double>(om, m, C);
    }
    return 0;
}

}  // extern "C"
