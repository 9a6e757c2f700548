// Stress-restricted dimension wrappers around J2 plasticity as a per-point STATE MACHINE (reference src/wrappers.jl:45-79 around
// src/Plastic.jl:126-172) — the form the lane-persistent kernel of mm_wrap.cu drives.
//
// One thread per point runs, per warp, the MAXIMUM of its lanes' work: the outer Newton on the out-of-plane strains takes 2-6 inner J2
// solves per point and every inner solve 0-3 Newton passes, so only 18.8 of 32 lanes were active per instruction
// (profiles/r02_wrapped_planestress_before_ncu.txt).  Here the same arithmetic is cut at its natural seams into four sections, and a lane
// carries a point through them independently of its neighbours:
//     TRIAL   σ_trial, Φ at the current 3D strain                                   Plastic.jl:131-135
//     NEWTON  ONE pass of the inner Newton (residuals, convergence test, closed-form step)  Plastic.jl:142-172
//     OUTER   out-of-plane stress norm, convergence test, outer update of the strain        wrappers.jl:60-63, :72-76
//     FINAL   Schur-complement tangent, reduced outputs, new state                          wrappers.jl:63-70
// A lane may pass through several sections in one trip of the kernel's loop (TRIAL -> NEWTON, NEWTON converged -> OUTER -> FINAL).
// The expressions are those of plastic_point_parts / wrapped_core, in the same order: the iterates, the inner and outer iteration
// counts and the flags are the same as the one-thread-per-point path produces (tests compare the two and both against the oracle).
// The old state, σ_trial and the three tangent ingredients live in the lane's scratch column (shared memory on the device), the Newton
// iterate and the 3D strain in registers.
#pragma once
#include "mm_wrap_math.cuh"

namespace mm {

enum { J2P_IDLE = 0, J2P_TRIAL = 1, J2P_NEWTON = 2, J2P_OUTER = 3, J2P_FINAL = 4 };
// scratch column: old state [εᵖ(6), κ, α(6), μ], σ_trial(6), tangent ingredients n(6), u(6), gξ
enum { J2S_EPN = 0, J2S_KAPN = 6, J2S_ALN = 7, J2S_MUN = 13, J2S_ST = 14, J2S_N = 20, J2S_U = 26, J2S_GXI = 32, J2S_SLOTS = 33 };

struct J2Lane {
    double eps[6];                    // 3D strain of the current outer iterate
    double sg[6], al[6], kap, mu;     // inner Newton iterate; after the inner solve: σ, α, κ, μ of this evaluation
    int iters;                        // Newton updates of the current inner solve
    int it;                           // outer evaluation number (1-based, wrappers.jl:60)
    int oit;                          // outer updates reported
    int flag;                         // inner flag bits (1 = yielded, 0x80 = inner failed) | 0x40 outer failed
    int phase;
    bool done;                        // outer Newton converged
};

// ---- load a point: reduced strain -> 3D strain with zero out-of-plane entries (increase_dim, wrappers.jl:24-27), old state -> scratch
template <int KIND, class Acc, class Scr> MM_HD void j2lp_load(Acc& acc, const Scr& scr, J2Lane& L) {
    constexpr int d = wrap_dim(KIND), nc = wrap_nc(KIND);
#pragma unroll
    for (int c = 0; c < 6; ++c) L.eps[c] = 0.0;
#pragma unroll
    for (int r = 0; r < nc; ++r) L.eps[wrap_slot(d, r)] = acc.template in<0>(r);
#pragma unroll
    for (int c = 0; c < 14; ++c) scr(J2S_EPN + c) = acc.template in<1>(c);
    L.it = 1; L.oit = 0; L.flag = 0; L.iters = 0; L.done = false;
    L.phase = J2P_TRIAL;
}

// ---- TRIAL                                                                           Plastic.jl:131-135
template <class Scr> MM_HD void j2lp_trial(const PlasticK& k, const Scr& scr, J2Lane& L) {
    const double c1 = 1.2247448713915890491;
    const double G2 = 2.0 * k.G;
    double e[6], st[6], d6[6], s[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) e[c] = L.eps[c] - scr(J2S_EPN + c);
    const double ltr = k.lam * tr6(e);
#pragma unroll
    for (int c = 0; c < 6; ++c) { st[c] = is_diag_slot(c) ? (G2 * e[c] + ltr) : (G2 * e[c]); scr(J2S_ST + c) = st[c]; }
#pragma unroll
    for (int c = 0; c < 6; ++c) { L.al[c] = scr(J2S_ALN + c); d6[c] = st[c] - L.al[c]; L.sg[c] = st[c]; }
    dev6(d6, s);
    double rho;
    fast_rsqrt(ddot6(s, s), &rho);
    L.kap = scr(J2S_KAPN);
    L.mu = scr(J2S_MUN);
    const double Phi = c1 * rho - k.sigma_y - L.kap;
    L.iters = 0;
    if (!(Phi <= 0.0)) {
        L.flag = 1;
        L.phase = J2P_NEWTON;
    } else {                                                   // elastic evaluation: σ_trial, Eᵉ, state unchanged
        L.flag = 0;
#pragma unroll
        for (int c = 0; c < 6; ++c) { scr(J2S_N + c) = 0.0; scr(J2S_U + c) = 0.0; }
        scr(J2S_GXI) = 0.0;
        L.phase = J2P_OUTER;
    }
}

// ---- NEWTON: one pass of the loop of plastic_point_parts (mm_math.cuh), same expressions in the same order
template <class Scr> MM_HD void j2lp_pass(const PlasticK& k, const Scr& scr, J2Lane& L) {
    const double c1 = 1.2247448713915890491;
    const double r2 = 1.4142135623730950488;
    const double G2 = 2.0 * k.G;
    double d6[6], s[6], n[6], w[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) d6[c] = L.sg[c] - L.al[c];
    dev6(d6, s);
    double rho;
    const double inv_rho = fast_rsqrt(ddot6(s, s), &rho);
#pragma unroll
    for (int c = 0; c < 6; ++c) n[c] = s[c] * inv_rho;
    dev6(L.al, w);
    const double kap_n = scr(J2S_KAPN);
    const double kr = L.kap * k.ikinf - 1.0;
    double Rs[6], Ra[6];
    const double gm = G2 * L.mu * c1, hm = L.mu * k.H_alpha;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        Rs[c] = (L.sg[c] - scr(J2S_ST + c)) + gm * n[c];
        Ra[c] = (L.al[c] - scr(J2S_ALN + c)) - hm * (k.c2 * w[c] - c1 * n[c]);
    }
    const double Rk = (L.kap - kap_n) - L.mu * k.H_kappa * kr;
    const double Rm = c1 * rho - k.sigma_y - L.kap;
    bool conv = (fabs(Rk) <= k.tol) & (fabs(Rm) <= k.tol);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const double t = is_diag_slot(c) ? k.tol : k.tol * (1.0 / r2);
        conv = conv & (fabs(Rs[c]) <= t) & (fabs(Ra[c]) <= t);
    }
    const double beta = c1 * inv_rho;
    const double a1 = 1.0 - hm * k.c2;
    const double a2 = 1.0 - L.mu * k.hk_ikinf;
    const double ia = fast_rcp(a1 * a2);
    const double i1psi = a2 * ia;
    const double ikk = a1 * ia;
    const double theta = gm * inv_rho;
    const double phi = hm * beta;
    const double eta1 = 1.0 + (theta - phi * i1psi);
    const double gk = -k.H_kappa * kr;
    double gt[6], ha[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        ha[c] = k.H_alpha * (c1 * n[c] - k.c2 * w[c]);
        gt[c] = G2 * c1 * n[c] - ha[c] * i1psi;
    }
    const double n_gt = ddot6(n, gt);
    const double den = c1 * n_gt - gk * ikk;
    const double ib = fast_rcp(den * eta1);
    const double i1eta = den * ib;
    const double iden = eta1 * ib;
    const bool fail = !conv && (L.iters >= k.max_iter || !(fabs(Rm) < 1.7e308));
    if (conv || fail) {
        // results at this iterate: tangent ingredients of ∂σ∂ε = Eᵉ − gξ I_dev + u ⊗ n                     Plastic.jl:150-155
        const double xi = theta * i1eta;
        const double omega = k.H_alpha * k.c2 * i1psi;
        const double cD = c1 * iden;
        const double nw = ddot6(n, w);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            scr(J2S_U + c) = G2 * (xi * n[c] + cD * (xi * omega * (w[c] - nw * n[c]) - G2 * c1 * n[c]));
            scr(J2S_N + c) = n[c];
        }
        scr(J2S_GXI) = G2 * xi;
        if (fail) { L.flag |= 0x80; L.phase = J2P_FINAL; }     // wrappers: the outer loop stops on a failed inner solve
        else L.phase = J2P_OUTER;
        return;
    }
    const double trRs = tr6(Rs) * (1.0 / 3.0), trRa = tr6(Ra) * (1.0 / 3.0);
    double rt[6], ra[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const double dl = is_diag_slot(c) ? 1.0 : 0.0;
        ra[c] = Ra[c] - trRa * dl;
        rt[c] = -(Rs[c] - trRs * dl) + ra[c] * i1psi;
    }
    const double n_rt = ddot6(n, rt);
    const double dmu = (c1 * n_rt + Rm + Rk * ikk) * iden;
    const double dkap = (-Rk - gk * dmu) * ikk;
    const double ny = n_rt - n_gt * dmu;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const double dl = is_diag_slot(c) ? 1.0 : 0.0;
        const double Pd = ((rt[c] - gt[c] * dmu) - ny * n[c]) * i1eta;
        L.sg[c] += -Rs[c] - theta * Pd - G2 * c1 * n[c] * dmu;
        L.al[c] += (-ra[c] - phi * Pd - ha[c] * dmu) * i1psi - trRa * dl;
    }
    L.kap += dkap;
    L.mu += dmu;
    ++L.iters;
}

template <class Scr> MM_HD double j2lp_T(const PlasticK& k, const Scr& scr, int I, int J) {
    return fma(scr(J2S_U + I), scr(J2S_N + J), fma(-scr(J2S_GXI), idev_entry(I, J), iso_entry(I, J, k.lam, k.G)));
}

// ---- OUTER: convergence test on the out-of-plane stresses, otherwise one outer Newton update       wrappers.jl:60-63, :72-77
template <class Scr> struct J2LaneTangent {
    const PlasticK& k; const Scr& scr;
    MM_HD double operator()(int I, int J) const { return j2lp_T(k, scr, I, J); }
};
template <int KIND, class Scr> MM_HD void j2lp_outer(const PlasticK& k, const Scr& scr, J2Lane& L, double tol, int max_iter) {
    L.oit = L.it - 1;
    const int r = wrap_outer_step<KIND>(J2LaneTangent<Scr>{k, scr}, L.sg, L.eps, tol);
    if (r == 0) { L.done = true; L.phase = J2P_FINAL; return; }
    if (r == 2) { L.flag |= 0x40; L.phase = J2P_FINAL; return; }
    L.oit = L.it;
    ++L.it;
    if (L.it > max_iter) { L.flag |= 0x40; L.phase = J2P_FINAL; }         // "Could not find plane stress state" (wrappers.jl:78)
    else L.phase = J2P_TRIAL;
}

// ---- FINAL: reduced stress, Schur-complement tangent, new state                                    wrappers.jl:63-70
// acc: out<0> reduced stress, out<1> reduced tangent (optional), out<2> new state
template <int KIND, class Acc, class Scr> MM_HD bool j2lp_final(const PlasticK& k, const Scr& scr, const J2Lane& L, Acc& acc) {
    constexpr int d = wrap_dim(KIND), nc = wrap_nc(KIND);
    const double c1 = 1.2247448713915890491;
#pragma unroll
    for (int r = 0; r < nc; ++r) acc.template out<0>(r, L.sg[wrap_slot(d, r)]);
    bool ok = true;
    if (L.done) ok = wrap_schur_out<KIND>(J2LaneTangent<Scr>{k, scr}, acc);
    else if (acc.template has_out<1>()) {
#pragma unroll
        for (int c = 0; c < nc * nc; ++c) acc.template out<1>(c, 0.0);
    }
    // new state of the last evaluation: εᵖ = εᵖₙ + μ ν̂ (yielded evaluations only), κ, α, μ                Plastic.jl:150-151
    const double mnu = L.mu * c1;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        double ep = scr(J2S_EPN + c);
        if (L.flag & 1) ep += mnu * scr(J2S_N + c);
        acc.template out<2>(c, ep);
    }
    acc.template out<2>(6, L.kap);
#pragma unroll
    for (int c = 0; c < 6; ++c) acc.template out<2>(7 + c, L.al[c]);
    acc.template out<2>(13, L.mu);
    return ok;
}

}  // namespace mm
