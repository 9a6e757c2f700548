// CrystalViscoPlastic per-point update (FP64) — reference src/CrystalViscoPlastic/CrystalViscoPlastic.jl:64-132
// with the package's own Newton `solve` (src/nonlinear_solver.jl:52-62): evaluate r and J, stop when
// ||r||_2 < tol, else x -= J \ r, at most max_iter loop passes.
//
// in<0> = Δε(6), in<1> = state [σ(6), κ(12), α(12), μ(12)];  out<0> = σ, out<1> = dσdε(36, optional), out<2> = new state
//
// What is restated analytically instead of by dual numbers (exact same derivatives):
//   τ_i = σ:MS_i,  σ = σ_trial − Σ_j μ_j s_j Eᵉ:MS_j      ⇒ ∂τ_i/∂μ_j = −s_j G_ij,  G_ij = (Eᵉ:MS_j):MS_i   (table)
//   κ_i = κₙ_i + Σ_j H[j,i] μ_j                            ⇒ ∂κ_i/∂μ_j = H[j,i]
//   α_i = (αₙ_i + H_kin μ_i s_i)/(1 + H_kin μ_i/α_∞)       ⇒ ∂α_i/∂μ_i = (H_kin s_i − α_i H_kin/α_∞)/(1 + H_kin μ_i/α_∞)
//   R_i = Δt (⟨Φ_i⟩/σ_c)^m − t* μ_i,  Φ_i = |τ_i − α_i| − τ_y − κ_i,  s_j = sign of the TRIAL reduced stress (frozen)
//   ∂R_i/∂Δε = p_i sgn(τ_i − α_i) Eᵉ:MS_i   with p_i = Δt m/σ_c (⟨Φ_i⟩/σ_c)^(m−1)
// Tangent (:79-89): dσdε = Eᵉ + Σ_i sgn_i (Eᵉ:MS_i) ⊗ z_i,  J Z = diag(p sgn) [Eᵉ:MS]  (6 right-hand sides through
// the LU of the converged Jacobian instead of forming inv(J)).
#pragma once
#include "mm_math.cuh"

namespace mm {

struct CrystalK {
    double lam, G;                    // Lamé constants of Eᵉ
    double tau_y, H_kin, alpha_inf, t_star, sigma_c, m, dt, tol;
    int max_iter, pad_;
    double Pw[12][6];                 // τ_i = Σ_c σ_c Pw[i][c]   (symmetric-storage weights of MS_i)
    double EM[12][6];                 // Eᵉ:MS_i, symmetric storage
    double Gm[12][12];                // Gm[i][j] = (Eᵉ:MS_j):MS_i
    double H[12][12];                 // H[j][i] = Julia H[j+1,i+1]
};

// host-side table build (once per material)
inline void make_crystal_k(CrystalK& k, double E, double nu, double tau_y, double H_kin, double alpha_inf, double t_star, double sigma_c,
                           double m, const double MS[12][9], const double H[12][12], double dt, double tol, int max_iter) {
    k.lam = E * nu / ((1 + nu) * (1 - 2 * nu));
    k.G = E / (2 * (1 + nu));
    k.tau_y = tau_y; k.H_kin = H_kin; k.alpha_inf = alpha_inf; k.t_star = t_star; k.sigma_c = sigma_c; k.m = m; k.dt = dt; k.tol = tol;
    k.max_iter = max_iter; k.pad_ = 0;
    static const int pi[6] = {0, 1, 2, 1, 2, 2}, pj[6] = {0, 0, 0, 1, 1, 2};
    for (int s = 0; s < 12; ++s) {
        const double* M = MS[s];    // column-major: M(i,j) = M[i + 3 j]
        const double trM = M[0] + M[4] + M[8];
        for (int c = 0; c < 6; ++c) {
            const int i = pi[c], j = pj[c];
            k.Pw[s][c] = (i == j) ? M[i + 3 * j] : (M[i + 3 * j] + M[j + 3 * i]);
            k.EM[s][c] = k.G * (M[i + 3 * j] + M[j + 3 * i]) + ((i == j) ? k.lam * trM : 0.0);
        }
    }
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j) {
            double g = 0.0;
            for (int c = 0; c < 6; ++c) g += k.EM[j][c] * k.Pw[i][c];
            k.Gm[i][j] = g;
            k.H[i][j] = H[i][j];
        }
}

// LU with partial pivoting on a 12x12 row-major matrix held in (thread-)local memory
MM_HD bool lu12_factor(double* A, int* piv) {
    for (int k = 0; k < 12; ++k) {
        int p = k; double amax = fabs(A[k * 12 + k]);
        for (int i = k + 1; i < 12; ++i) { const double v = fabs(A[i * 12 + k]); if (v > amax) { amax = v; p = i; } }
        piv[k] = p;
        if (amax == 0.0) return false;
        if (p != k) for (int j = 0; j < 12; ++j) { const double t = A[k * 12 + j]; A[k * 12 + j] = A[p * 12 + j]; A[p * 12 + j] = t; }
        const double ip = 1.0 / A[k * 12 + k];
        for (int i = k + 1; i < 12; ++i) {
            const double l = A[i * 12 + k] * ip;
            A[i * 12 + k] = l;
            for (int j = k + 1; j < 12; ++j) A[i * 12 + j] -= l * A[k * 12 + j];
        }
    }
    return true;
}
MM_HD void lu12_solve(const double* LU, const int* piv, double* b) {
    for (int k = 0; k < 12; ++k) { const int p = piv[k]; if (p != k) { const double t = b[k]; b[k] = b[p]; b[p] = t; } }
    for (int i = 1; i < 12; ++i) { double s = b[i]; for (int j = 0; j < i; ++j) s -= LU[i * 12 + j] * b[j]; b[i] = s; }
    for (int i = 11; i >= 0; --i) { double s = b[i]; for (int j = i + 1; j < 12; ++j) s -= LU[i * 12 + j] * b[j]; b[i] = s / LU[i * 12 + i]; }
}

// returns the active mask (bit i <=> Φ_i > 0 at the solution; bit 15 = not converged); *iters_out = number of x updates
template <class Acc> MM_HD unsigned crystal_point(const CrystalK& k, Acc& acc, int* iters_out) {
    const double G2 = 2.0 * k.G;
    double st[6];
    {   // σ_trial = σₙ + Eᵉ ⊡ Δε                                                           :119
        double e[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) e[c] = acc.template in<0>(c);
        const double ltr = k.lam * tr6(e);
#pragma unroll
        for (int c = 0; c < 6; ++c) st[c] = acc.template in<1>(c) + (is_diag_slot(c) ? (G2 * e[c] + ltr) : (G2 * e[c]));
    }
    double kap_n[12], al_n[12], x[12], sg[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) { kap_n[i] = acc.template in<1>(6 + i); al_n[i] = acc.template in<1>(18 + i); x[i] = acc.template in<1>(30 + i); }
#pragma unroll
    for (int i = 0; i < 12; ++i) {                                                      // :120-121
        double tau = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) tau += st[c] * k.Pw[i][c];
        const double d = tau - al_n[i];
        sg[i] = (d > 0.0) ? 1.0 : ((d < 0.0) ? -1.0 : 0.0);
    }
    double J[144], sig[6], kap[12], al[12], p[12], sgn[12];
    int piv[12];
    unsigned mask = 0;
    int iters = 0;
    bool converged = false;
    for (int it = 1;; ++it) {   // pass max_iter+1 only re-evaluates the state at the last iterate of a failed solve
        // state variables at x                                                            :118-132
#pragma unroll
        for (int c = 0; c < 6; ++c) sig[c] = st[c];
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const double f = x[j] * sg[j];
#pragma unroll
            for (int c = 0; c < 6; ++c) sig[c] -= f * k.EM[j][c];
        }
        double r[12], nrm2 = 0.0;
        mask = 0;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            double kp = 0.0;
#pragma unroll
            for (int j = 0; j < 12; ++j) kp += k.H[j][i] * x[j];
            kap[i] = kap_n[i] + kp;
            const double den = 1.0 + k.H_kin * x[i] / k.alpha_inf;
            al[i] = (al_n[i] + k.H_kin * x[i] * sg[i]) / den;
            double tau = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) tau += sig[c] * k.Pw[i][c];
            const double red = tau - al[i];
            sgn[i] = (red > 0.0) ? 1.0 : ((red < 0.0) ? -1.0 : 0.0);
            const double Phi = fabs(red) - k.tau_y - kap[i];                            // :110
            double Ri = -k.t_star * x[i];
            p[i] = 0.0;
            if (Phi > 0.0) {                                                            // ⟨Φ⟩ = (Φ+|Φ|)/2
                const double xx = Phi / k.sigma_c;
                const double pw1 = pow(xx, k.m - 1.0);
                Ri = k.dt * (pw1 * xx) - k.t_star * x[i];                               // :111
                p[i] = k.dt * k.m / k.sigma_c * pw1;
                mask |= (1u << i);
            }
            r[i] = Ri;
            nrm2 += Ri * Ri;
            const double dal = (k.H_kin * sg[i] - al[i] * k.H_kin / k.alpha_inf) / den;
            // Jacobian row i
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                double v = p[i] * (sgn[i] * (-sg[j] * k.Gm[i][j]) - k.H[j][i]);
                if (j == i) v += -p[i] * sgn[i] * dal - k.t_star;
                J[i * 12 + j] = v;
            }
        }
        iters = it - 1;
        if (it > k.max_iter) break;                                                      // nonlinear_solver.jl:61 (returns nothing)
        if (sqrt(nrm2) < k.tol) { converged = true; break; }                             // nonlinear_solver.jl:56
        if (!lu12_factor(J, piv)) break;
        lu12_solve(J, piv, r);
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] -= r[i];                                      // nonlinear_solver.jl:59
        iters = it;
    }
    if (!converged) mask |= 0x8000u;
    // outputs                                                                            :77, :90
#pragma unroll
    for (int c = 0; c < 6; ++c) { acc.template out<0>(c, sig[c]); acc.template out<2>(c, sig[c]); }
#pragma unroll
    for (int i = 0; i < 12; ++i) { acc.template out<2>(6 + i, kap[i]); acc.template out<2>(18 + i, al[i]); acc.template out<2>(30 + i, x[i]); }
    if (acc.template has_out<1>()) {
        double tan[36];
#pragma unroll
        for (int Jc = 0; Jc < 6; ++Jc)
#pragma unroll
            for (int Ic = 0; Ic < 6; ++Ic) tan[Ic + 6 * Jc] = iso_entry(Ic, Jc, k.lam, k.G);
        if ((mask & 0x0FFFu) != 0 && converged && lu12_factor(J, piv)) {                 // :79-89
            for (int Jc = 0; Jc < 6; ++Jc) {
                double z[12];
                for (int j = 0; j < 12; ++j) z[j] = p[j] * sgn[j] * k.EM[j][Jc];
                lu12_solve(J, piv, z);
                for (int i = 0; i < 12; ++i) {
                    const double zi = sgn[i] * z[i];
                    for (int Ic = 0; Ic < 6; ++Ic) tan[Ic + 6 * Jc] += k.EM[i][Ic] * zi;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 36; ++c) acc.template out<1>(c, tan[c]);
    }
    *iters_out = iters;
    return mask;
}

}  // namespace mm
