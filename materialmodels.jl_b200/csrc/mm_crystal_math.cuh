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
    // structured fast path (crystal_point_fast): H = a_h I + b_h 11ᵀ as the reference constructor builds it (:41)
    double a_h, b_h;
    double Cmp[6][6];                 // (Eᵉ)⁻¹ as a 6x6 on storage components (raw inverse of Ee[I][J])
    double isc, dtm_sc, hk_ai;        // 1/σ_c, Δt m/σ_c, H_kin/α_∞
    double mits, tol2_lo, tol2_hi;    // −1/t*, tol²(1 ∓ 8 eps)
    int h_struct;                     // 1 if H has the form above (else only the general path applies)
    int m_int;                        // m if it is an integer in [2,64] (x^(m-1) by multiplications), else 0 (pow)
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
    const double hd = H[0][0], ho = H[0][1];
    k.h_struct = 1;
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j)
            if (H[i][j] != ((i == j) ? hd : ho)) k.h_struct = 0;
    k.b_h = ho; k.a_h = hd - ho;
    {   // inverse of the raw 6x6 [λ 11ᵀ + 2G I on the normal block, G on the shear diagonal]
        const double i2G = 1.0 / (2.0 * k.G), f = k.lam / (3.0 * k.lam + 2.0 * k.G);
        for (int a = 0; a < 6; ++a)
            for (int b = 0; b < 6; ++b) {
                const bool da = (a == 0 || a == 3 || a == 5), db = (b == 0 || b == 3 || b == 5);
                k.Cmp[a][b] = (da && db) ? i2G * ((a == b ? 1.0 : 0.0) - f) : ((a == b) ? 1.0 / k.G : 0.0);
            }
    }
    k.isc = 1.0 / sigma_c; k.dtm_sc = dt * m / sigma_c; k.hk_ai = H_kin / alpha_inf;
    k.mits = -1.0 / t_star; k.tol2_lo = tol * tol * (1.0 - 1.8e-15); k.tol2_hi = tol * tol * (1.0 + 1.8e-15);
    k.m_int = (m >= 2.0 && m <= 64.0 && m == (double)(int)m) ? (int)m : 0;
}

// LU with partial pivoting on a 12x12 row-major matrix; element e lives at A[e * st] (st = 1: thread-local array; st = threads per
// block: the thread's column of a shared-memory image — the 144-entry matrix is what makes the general path slow when it sits in
// local memory, ~1 700 accesses per Newton pass at L2 latency once the per-thread stacks overflow L1)
MM_HD bool lu12_factor(double* A, int* piv, int st = 1) {
    for (int k = 0; k < 12; ++k) {
        int p = k; double amax = fabs(A[(k * 12 + k) * st]);
        for (int i = k + 1; i < 12; ++i) { const double v = fabs(A[(i * 12 + k) * st]); if (v > amax) { amax = v; p = i; } }
        piv[k] = p;
        if (amax == 0.0) return false;
        if (p != k) for (int j = 0; j < 12; ++j) { const double t = A[(k * 12 + j) * st]; A[(k * 12 + j) * st] = A[(p * 12 + j) * st]; A[(p * 12 + j) * st] = t; }
        const double ip = 1.0 / A[(k * 12 + k) * st];
        for (int i = k + 1; i < 12; ++i) {
            const double l = A[(i * 12 + k) * st] * ip;
            A[(i * 12 + k) * st] = l;
            for (int j = k + 1; j < 12; ++j) A[(i * 12 + j) * st] -= l * A[(k * 12 + j) * st];
        }
    }
    return true;
}
MM_HD void lu12_solve(const double* LU, const int* piv, double* b, int st = 1) {
    for (int k = 0; k < 12; ++k) { const int p = piv[k]; if (p != k) { const double t = b[k]; b[k] = b[p]; b[p] = t; } }
    for (int i = 1; i < 12; ++i) { double s = b[i]; for (int j = 0; j < i; ++j) s -= LU[(i * 12 + j) * st] * b[j]; b[i] = s; }
    for (int i = 11; i >= 0; --i) { double s = b[i]; for (int j = i + 1; j < 12; ++j) s -= LU[(i * 12 + j) * st] * b[j]; b[i] = s / LU[(i * 12 + i) * st]; }
}

// General path in resumable form (init / one `solve` pass / finalize), so that a kernel can interleave points lane by lane
// (the fix-up pass of mm_crystal.cu refills a lane as soon as its solve ends; with one monolithic solve per thread a single
// non-converging point holds its whole warp for max_iter passes).  crystal_point below is init + passes + finalize.
struct CrystalGen {
    double st[6], kap_n[12], al_n[12], x[12], sg[12];      // inputs of the solve and the iterate
    double sig[6], kap[12], al[12], p[12], sgn[12];        // state variables at the last evaluation
    double* J;                                             // Jacobian at the last evaluation (factored in place by the update): element e at J[e * Js]
    int Js;
    unsigned mask;
    int it, iters;
    bool converged;
};

template <class Acc> MM_HD void crystal_gen_init(const CrystalK& k, Acc& acc, CrystalGen& S, double* J, int Js) {
    const double G2 = 2.0 * k.G;
    S.J = J; S.Js = Js;
    {   // σ_trial = σₙ + Eᵉ ⊡ Δε                                                           :119
        double e[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) e[c] = acc.template in<0>(c);
        const double ltr = k.lam * tr6(e);
#pragma unroll
        for (int c = 0; c < 6; ++c) S.st[c] = acc.template in<1>(c) + (is_diag_slot(c) ? (G2 * e[c] + ltr) : (G2 * e[c]));
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) { S.kap_n[i] = acc.template in<1>(6 + i); S.al_n[i] = acc.template in<1>(18 + i); S.x[i] = acc.template in<1>(30 + i); }
#pragma unroll
    for (int i = 0; i < 12; ++i) {                                                      // :120-121
        double tau = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) tau += S.st[c] * k.Pw[i][c];
        const double d = tau - S.al_n[i];
        S.sg[i] = (d > 0.0) ? 1.0 : ((d < 0.0) ? -1.0 : 0.0);
    }
    S.mask = 0; S.it = 1; S.iters = 0; S.converged = false;
}

// one pass of the `solve` loop (nonlinear_solver.jl:53-60): evaluate r and J at x, test, update.  true = finished.
// (pass max_iter+1 only re-evaluates the state at the last iterate of a failed solve)
MM_HD bool crystal_gen_pass(const CrystalK& k, CrystalGen& S) {
    // state variables at x                                                            :118-132
#pragma unroll
    for (int c = 0; c < 6; ++c) S.sig[c] = S.st[c];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        const double f = S.x[j] * S.sg[j];
#pragma unroll
        for (int c = 0; c < 6; ++c) S.sig[c] -= f * k.EM[j][c];
    }
    double r[12], nrm2 = 0.0;
    unsigned mask = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        double kp = 0.0;
#pragma unroll
        for (int j = 0; j < 12; ++j) kp += k.H[j][i] * S.x[j];
        S.kap[i] = S.kap_n[i] + kp;
        const double den = 1.0 + k.H_kin * S.x[i] / k.alpha_inf;
        S.al[i] = (S.al_n[i] + k.H_kin * S.x[i] * S.sg[i]) / den;
        double tau = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) tau += S.sig[c] * k.Pw[i][c];
        const double red = tau - S.al[i];
        S.sgn[i] = (red > 0.0) ? 1.0 : ((red < 0.0) ? -1.0 : 0.0);
        const double Phi = fabs(red) - k.tau_y - S.kap[i];                              // :110
        double Ri = -k.t_star * S.x[i];
        S.p[i] = 0.0;
        if (!(Phi <= 0.0) && !(Phi > 0.0)) Ri = Phi;                                    // NaN Φ: ⟨Φ⟩ = (Φ+|Φ|)/2 is NaN in the reference, so is R
        if (Phi > 0.0) {                                                                // ⟨Φ⟩ = (Φ+|Φ|)/2
            const double xx = Phi / k.sigma_c;
            const double pw1 = pow(xx, k.m - 1.0);
            Ri = k.dt * (pw1 * xx) - k.t_star * S.x[i];                                 // :111
            S.p[i] = k.dt * k.m / k.sigma_c * pw1;
            mask |= (1u << i);
        }
        r[i] = Ri;
        nrm2 += Ri * Ri;
        const double dal = (k.H_kin * S.sg[i] - S.al[i] * k.H_kin / k.alpha_inf) / den;
        // Jacobian row i
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            double v = S.p[i] * (S.sgn[i] * (-S.sg[j] * k.Gm[i][j]) - k.H[j][i]);
            if (j == i) v += -S.p[i] * S.sgn[i] * dal - k.t_star;
            S.J[(i * 12 + j) * S.Js] = v;
        }
    }
    S.mask = mask;
    S.iters = S.it - 1;
    if (S.it > k.max_iter) return true;                                                  // nonlinear_solver.jl:61 (returns nothing)
    if (!(nrm2 == nrm2)) return true;                                                    // NaN residual can never pass :56 — the reference runs out of iterations
    if (sqrt(nrm2) < k.tol) { S.converged = true; return true; }                         // nonlinear_solver.jl:56
    int piv[12];
    if (!lu12_factor(S.J, piv, S.Js)) return true;
    lu12_solve(S.J, piv, r, S.Js);
#pragma unroll
    for (int i = 0; i < 12; ++i) S.x[i] -= r[i];                                        // nonlinear_solver.jl:59
    S.iters = S.it;
    S.it += 1;
    return false;
}

// outputs (:77, :90) and the tangent (:79-89); returns the active mask (bit i <=> Φ_i > 0 at the solution; bit 15 = not converged)
template <class Acc> MM_HD unsigned crystal_gen_finalize(const CrystalK& k, Acc& acc, CrystalGen& S) {
    unsigned mask = S.mask;
    if (!S.converged) mask |= 0x8000u;
#pragma unroll
    for (int c = 0; c < 6; ++c) { acc.template out<0>(c, S.sig[c]); acc.template out<2>(c, S.sig[c]); }
#pragma unroll
    for (int i = 0; i < 12; ++i) { acc.template out<2>(6 + i, S.kap[i]); acc.template out<2>(18 + i, S.al[i]); acc.template out<2>(30 + i, S.x[i]); }
    if (acc.template has_out<1>()) {
        double tan[36];
        int piv[12];
#pragma unroll
        for (int Jc = 0; Jc < 6; ++Jc)
#pragma unroll
            for (int Ic = 0; Ic < 6; ++Ic) tan[Ic + 6 * Jc] = iso_entry(Ic, Jc, k.lam, k.G);
        if ((mask & 0x0FFFu) != 0 && S.converged && lu12_factor(S.J, piv, S.Js)) {       // :79-89
            for (int Jc = 0; Jc < 6; ++Jc) {
                double z[12];
                for (int j = 0; j < 12; ++j) z[j] = S.p[j] * S.sgn[j] * k.EM[j][Jc];
                lu12_solve(S.J, piv, z, S.Js);
                for (int i = 0; i < 12; ++i) {
                    const double zi = S.sgn[i] * z[i];
                    for (int Ic = 0; Ic < 6; ++Ic) tan[Ic + 6 * Jc] += k.EM[i][Ic] * zi;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 36; ++c) acc.template out<1>(c, tan[c]);
    }
    return mask;
}

// returns the active mask (bit i <=> Φ_i > 0 at the solution; bit 15 = not converged); *iters_out = number of x updates
// J / Js: where the 144-entry Jacobian lives (NULL: a thread-local array)
template <class Acc> MM_HD unsigned crystal_point(const CrystalK& k, Acc& acc, int* iters_out, double* J = nullptr, int Js = 1) {
    CrystalGen S;
    double Jloc[144];
    crystal_gen_init(k, acc, S, J ? J : Jloc, J ? Js : 1);
    while (!crystal_gen_pass(k, S)) {}
    *iters_out = S.iters;
    return crystal_gen_finalize(k, acc, S);
}

// =================================================================================================
// Register-resident fast path.  Same Newton iterates as crystal_point (and as the reference), but the
// step J δ = r is solved through the structure of J instead of a 12x12 LU:
//     J = D − p̂ 1ᵀ − A Bᵀ ,   D_ii = −(t* + p_i (sgn_i α'_i + a_h)),  p̂ = b_h p   (cross hardening, rank 1)
//     A_i: = c_i Pw[i][:]  (c_i = p_i sgn_i),   B_j: = s_j EM[j][:] = s_j Eᵉ Pw[j][:]   (elastic coupling, rank 6)
// Woodbury:  (D − A Bᵀ)⁻¹ v = D⁻¹v + D⁻¹A K⁻¹ Bᵀ D⁻¹ v,  K = I₆ − Bᵀ D⁻¹ A = I₆ + Eᵉ S,  S = Σ_i w_i Pw_i Pw_iᵀ,
//            w_i = −s_i c_i / D_ii  (≥ 0 while the reduced stress keeps the sign of its trial value).
// K itself is non-symmetric (unpivoted LU of it loses 4-5 digits on the reference's non-traceless BCC systems —
// measured); multiplying K z = Bᵀ D⁻¹ v by the compliance (Eᵉ)⁻¹ gives the SYMMETRIC POSITIVE DEFINITE system
//            ((Eᵉ)⁻¹ + S) z = Σ_i s_i (D⁻¹v)_i Pw_i
// which an unpivoted LDLᵀ solves stably; a non-positive pivot (possible only when some w_i < 0) or a D_ii too close
// to zero sets *fallback and the caller re-runs the point through the general pivoted path.
// Sherman–Morrison adds the rank-1 cross-hardening term.  The only matrix is the 6x6 symmetric M (21 doubles) — it and
// the iterate live in registers with static indexing; the per-system vectors (κₙ, αₙ, c, D⁻¹r, 1/D) sit in the
// thread's shared-memory scratch column.  Requires k.h_struct.
// Scratch slots: [0,12) κₙ  [12,24) αₙ  [24,36) c  [36,48) D⁻¹r  [48,60) 1/D
// =================================================================================================
#define MM_CRYSTAL_SCRATCH 60

MM_HD double powi_(double b, int n) {      // b^n, n >= 0, by squaring (n is warp-uniform)
    double r = 1.0;
    while (n) { if (n & 1) r *= b; b *= b; n >>= 1; }
    return r;
}
// b^n for 1 <= n < 64 (n warp-uniform): same multiplication sequence as powi_, fixed trip count => straight-line code
MM_HD double powi6_(double b, int n) {
    double r = (n & 1) ? b : 1.0;
#pragma unroll
    for (int bit = 1; bit < 6; ++bit) {
#if defined(__CUDA_ARCH__)
        // predicated multiplies: n is uniform, so these are @P DMUL (no select pairs, no FP64 pipe time when off)
        asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p mul.f64 %0, %0, %1;\n\t}" : "+d"(b) : "d"(b), "r"(n >> bit));
        asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p mul.f64 %0, %0, %1;\n\t}" : "+d"(r) : "d"(b), "r"((n >> bit) & 1));
#else
        if ((n >> bit) != 0) b *= b;
        if ((n >> bit) & 1) r *= b;
#endif
    }
    return r;
}
// reciprocal without the range guard of fast_rcp: for arguments that are well scaled whenever the solve is healthy
// (1 + H_kin μ/α_∞ and the guarded D_ii); on a diverged iterate (0, inf, NaN) it yields inf/NaN like a division would,
// the residual norm turns NaN and the point ends as "not converged".
MM_HD double rcp_nc(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / x;
#endif
}

// in-place LDLᵀ of a symmetric 6x6 (lower triangle of a row-major array in registers), no pivoting.
// On exit: strict lower triangle = L, diagonal = 1/d_k.  false if a pivot is not safely positive.
MM_HD bool ldl6_factor(double* M) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double d = M[k * 6 + k];
        ok = ok & (d > 1e-200) & (d < 1e200);
        const double id = fast_rcp(d);
        M[k * 6 + k] = id;
        double col[6];
#pragma unroll
        for (int i = k + 1; i < 6; ++i) col[i] = M[i * 6 + k];          // unscaled column k
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            const double l = col[i] * id;
            M[i * 6 + k] = l;
#pragma unroll
            for (int j = k + 1; j <= i; ++j) M[i * 6 + j] -= l * col[j];
        }
    }
    return ok;
}
MM_HD void ldl6_solve(const double* M, double* b) {
#pragma unroll
    for (int i = 1; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) b[i] -= M[i * 6 + j] * b[j];
#pragma unroll
    for (int i = 0; i < 6; ++i) b[i] *= M[i * 6 + i];
#pragma unroll
    for (int i = 4; i >= 0; --i)
#pragma unroll
        for (int j = i + 1; j < 6; ++j) b[i] -= M[j * 6 + i] * b[j];
}

// ---- the structured path in three phases, so that a persistent warp can interleave points (mm_crystal.cu) ----
// Per-lane state: everything indexed by slip system lives in the scratch column (run-time indexed, rolled loops =>
// small code, no instruction-cache thrash); registers hold only the 6x6 M, σ and a few scalars.
// Scratch slots: [0,12) κₙ [12,24) αₙ [24,36) g = c/D = p·sgn/D_ii [36,48) D⁻¹r [48,60) x=μ [60,66) σ_trial
#undef MM_CRYSTAL_SCRATCH
#define MM_CRYSTAL_SCRATCH 66
constexpr int XS = 48, SIGT = 60;

struct CrystalLane {
    double K[36];          // lower triangle of M = (Eᵉ)⁻¹ + S at the last evaluation (kept for the tangent)
    double sig[6];         // σ at the last evaluation
    double yp[6];          // Σ s_i (D⁻¹p̂)_i Pw_i at the last evaluation (cross hardening only)
    unsigned spos, sneg;   // frozen trial signs
    unsigned mask;         // active mask at the last evaluation
    unsigned nz;           // systems with x_i != 0 (sparse pass: only these, and the systems that are active, cost more than a yield test)
    unsigned cand;         // sparse pass: candidate set of the last evaluation (their g / D⁻¹r slots are the only non-zero ones)
    int it;                // loop index of nonlinear_solver.jl:53 (1-based) of the NEXT evaluation
    int iters;             // number of x updates done
    bool converged, bad;
    bool big;              // active-set pass: the point left the small route (list longer than 6 / zero frozen sign) -> redo it through the sparse pass
    bool flip;             // some ACTIVE system has sgn_i != s_i at the last evaluation
};

MM_HD int __ffs_(unsigned m) {      // index of the lowest set bit (m != 0)
#if defined(__CUDA_ARCH__)
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}
// s_i = ±1.0 or 0.0 from the two frozen-sign bit masks, assembled in the integer pipe (high word 0x3FF00000 | sign bit; the FP64 select
// chain the ternary form compiles to was 4 % of the Newton kernel's instructions)
MM_HD double crystal_sign(const CrystalLane& L, int i) {
#if defined(__CUDA_ARCH__)
    const unsigned neg = (L.sneg >> i) & 1u, any = ((L.spos | L.sneg) >> i) & 1u;
    return __hiloint2double((int)((0u - any) & (0x3FF00000u | (neg << 31))), 0);
#else
    return ((L.spos >> i) & 1u) ? 1.0 : (((L.sneg >> i) & 1u) ? -1.0 : 0.0);
#endif
}
// sign(x) = ±1.0, 0.0 for x = 0 (Julia's sign), assembled from the sign bit in the integer pipe: one compare and one integer select
// instead of two FP64 compares and four FSEL
MM_HD double sign_of(double x) {
#if defined(__CUDA_ARCH__)
    const unsigned hi = (x == 0.0) ? 0u : (0x3FF00000u | ((unsigned)__double2hiint(x) & 0x80000000u));
    return __hiloint2double((int)hi, 0);
#else
    return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : 0.0);
#endif
}
// s_i · v for a system whose frozen sign is ±1 (every candidate of the active-set pass: a zero sign sends the point to the sparse launch first):
// the sign bit of v is flipped in the integer pipe instead of a DMUL with an assembled ±1.0
MM_HD double crystal_apply_sign(const CrystalLane& L, int i, double v) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(v) ^ (int)(((L.sneg >> i) & 1u) << 31), __double2loint(v));
#else
    return ((L.sneg >> i) & 1u) ? -v : v;
#endif
}
// b^n, n = NPOW known at compile time (the reference's m = 10 -> x^9: three squarings and one product, no predicates)
template <int NPOW> MM_HD double powc_(double b) {      // the multiplication sequence of powi6_ with the tests folded away
    double r = (NPOW & 1) ? b : 1.0;
    bool r_is_one = !(NPOW & 1);
#pragma unroll
    for (int bit = 1; bit < 7; ++bit) {
        if ((NPOW >> bit) != 0) b *= b;
        if ((NPOW >> bit) & 1) { r = r_is_one ? b : r * b; r_is_one = false; }       // 1.0 * b is exact, so skipping it changes nothing
    }
    return r;
}

// XSLOT: first scratch slot of x; SIGSLOT: first slot of σ_trial, or -1 = returned in st_out only (the compact layout of the active-set kernel)
// x^(m−1) of :111 in the per-lane-list passes (sparse, active-set).  Integer m: by multiplications (m = 10: straight-line x^9).  General m: pow() for
// the active systems; an inactive one carries x = ∓0 (or NaN), whose power is the same zero (or NaN).  INTPOW is a template parameter: the integer
// instantiation carries no pow() code (a run-time branch on m cost the m = 10 kernel 5 % more instructions).
template <bool INTPOW> MM_HD double crystal_pow_m1(const CrystalK& k, double xx, bool act) {
    if (INTPOW) { const int npow = k.m_int - 1; return (npow == 9) ? powc_<9>(xx) : powi6_(xx, npow); }
    return act ? pow(xx, k.m - 1.0) : xx;
}

template <int XSLOT, int SIGSLOT, class KT, class Acc, class Scr>       // KT: CrystalK or CrystalNK (reads lam, G, Pw only)
MM_HD void crystal_init_t(const KT& k, Acc& acc, Scr& scr, CrystalLane& L, double* st_out) {
    const double G2 = 2.0 * k.G;
    double st[6];
    {   // σ_trial = σₙ + Eᵉ ⊡ Δε                                                           :119
        double e[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) e[c] = acc.template in<0>(c);
        const double ltr = k.lam * tr6(e);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            st[c] = acc.template in<1>(c) + (is_diag_slot(c) ? (G2 * e[c] + ltr) : (G2 * e[c]));
            if (SIGSLOT >= 0) scr((SIGSLOT >= 0 ? SIGSLOT : 0) + c) = st[c];
            if (st_out) st_out[c] = st[c];
        }
    }
    L.spos = 0; L.sneg = 0;
    unsigned nz = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {                                                      // :120-121
        const double kn = acc.template in<1>(6 + i), an = acc.template in<1>(18 + i), x0 = acc.template in<1>(30 + i);
        scr(i) = kn; scr(12 + i) = an; scr(XSLOT + i) = x0;
        if (x0 != 0.0) nz |= (1u << i);
        double tau = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) tau += st[c] * k.Pw[i][c];
        const double d = tau - an;
        if (d > 0.0) L.spos |= (1u << i);
        if (d < 0.0) L.sneg |= (1u << i);
    }
    L.mask = 0; L.nz = nz; L.cand = 0xFFFu; L.it = 1; L.iters = 0; L.converged = false; L.bad = false; L.big = false; L.flip = false;
}
template <class KT, class Acc, class Scr>
MM_HD void crystal_init(const KT& k, Acc& acc, Scr& scr, CrystalLane& L) { crystal_init_t<XS, SIGT>(k, acc, scr, L, nullptr); }

#ifndef MM_CRYSTAL_UNROLL
#define MM_CRYSTAL_UNROLL 2   // r02 sweep after the sign / power specialisation: 1 -> 22.0 ms, 2 -> 21.5 ms, 3 -> 21.6 ms (profiles/r02_crystal_variant_and_knob_sweep.log)
#endif
#ifndef MM_CRYSTAL_UNROLL_SMALL
#define MM_CRYSTAL_UNROLL_SMALL 12
#endif
constexpr int kCrystalUnroll = MM_CRYSTAL_UNROLL, kCrystalUnrollSmall = MM_CRYSTAL_UNROLL_SMALL;

// Evaluate r and J at x (the first half of a `solve` pass, nonlinear_solver.jl:54-55): on exit L.sig = σ(x), L.K = M =
// (Eᵉ)⁻¹ + S (lower triangle), L.yp, L.mask, L.flip are set, the scratch holds c, D⁻¹r, 1/D, and yr = Σ s_i (D⁻¹r)_i Pw_i.
// INTPOW (integer m: x^(m-1) of :111 by multiplications): branch-free over the slip systems — inactive systems run the
// same arithmetic with ⟨Φ⟩ = 0 (p = 0, R = −t* μ), identical values, no divergence inside the pass, and the per-system
// dependency chains (two reciprocals, one power) can overlap.  !INTPOW: general m through pow(), active systems only.
template <bool HASQ, bool INTPOW, class Scr>
MM_HD void crystal_evaluate(const CrystalK& k, Scr& scr, CrystalLane& L, double* yr, double* nrm2_out, bool* bad_out) {
    double sumx = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) L.sig[c] = scr(SIGT + c);
#pragma unroll kCrystalUnrollSmall
    for (int j = 0; j < 12; ++j) {                                                      // σ = σ_trial − Σ x_j s_j Eᵉ:MS_j   :131
        const double xj = scr(XS + j);
        const double f = xj * crystal_sign(L, j);
        if (HASQ) sumx += xj;
#pragma unroll
        for (int c = 0; c < 6; ++c) L.sig[c] -= f * k.EM[j][c];
    }
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) L.K[a * 6 + b] = k.Cmp[a][b];
#pragma unroll
    for (int c = 0; c < 6; ++c) { yr[c] = 0.0; L.yp[c] = 0.0; }
    double nrm2 = 0.0;
    unsigned mask = 0;
    bool bad = false, flip = false;
    const int npow = k.m_int - 1;
#pragma unroll kCrystalUnroll
    for (int i = 0; i < 12; ++i) {
        double pw[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) pw[c] = k.Pw[i][c];
        const double si = crystal_sign(L, i), xi = scr(XS + i);
        double kap = scr(i) + k.a_h * xi;                                               // :127 with H = a I + b 11ᵀ
        if (HASQ) kap += k.b_h * sumx;
        const double iden = INTPOW ? rcp_nc(1.0 + k.hk_ai * xi) : fast_rcp(1.0 + k.hk_ai * xi);
        const double al = (scr(12 + i) + k.H_kin * xi * si) * iden;                     // :128
        double tau = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) tau += L.sig[c] * pw[c];
        const double red = tau - al;
        const double sgn = sign_of(red);
        const double Phi = fabs(red) - k.tau_y - kap;                                    // :110
        const bool act = Phi > 0.0;                                                     // ⟨Φ⟩ = (Φ+|Φ|)/2
        double Ri, p, invD;
        if (INTPOW) {
            const double xx = act ? Phi * k.isc : Phi * 0.0;                            // ∓0 when inactive; a NaN Φ stays NaN (⟨Φ⟩ = (Φ+|Φ|)/2 in the reference)
            const double pw1 = (npow == 9) ? powc_<9>(xx) : powi6_(xx, npow);     // m = 10 (the reference's value): straight-line x^9
            Ri = k.dt * (pw1 * xx) - k.t_star * xi;                                     // :111  (inactive: 0 − t* μ)
            p = k.dtm_sc * pw1;
            mask |= act ? (1u << i) : 0u;
            flip = flip | (act & (sgn != si));
            const double dal = (k.H_kin * si - al * k.hk_ai) * iden;
            // D_ii = −dD.  The Woodbury split needs D safely invertible: dD >= t*/4 (it is >= t* unless an ACTIVE system's
            // reduced stress has flipped sign against its trial value while a_h < H_kin); otherwise -> pivoted general path.
            const double dD = k.t_star + p * (sgn * dal + k.a_h);
            bad = bad | (act & !(dD >= 0.25 * k.t_star));
            invD = act ? -rcp_nc(dD) : k.mits;                                          // inactive: D_ii = −t*
        } else {
            Ri = -k.t_star * xi; p = 0.0; invD = k.mits;
            if (!act && !(Phi <= 0.0)) Ri = Phi;                                        // NaN Φ
            if (act) {
                const double xx = Phi * k.isc;
                const double pw1 = pow(xx, k.m - 1.0);
                Ri = k.dt * (pw1 * xx) - k.t_star * xi;
                p = k.dtm_sc * pw1;
                mask |= (1u << i);
                flip = flip | (sgn != si);
                const double dal = (k.H_kin * si - al * k.hk_ai) * iden;
                const double dD = k.t_star + p * (sgn * dal + k.a_h);
                bad = bad | !(dD >= 0.25 * k.t_star);
                invD = -fast_rcp(dD);
            }
        }
        nrm2 += Ri * Ri;
        const double ci = p * sgn;
        const double rD = Ri * invD;
        const double g = ci * invD;                                                     // invD < 0 always, so sign(c_i) = −sign(g_i)
        scr(24 + i) = g; scr(36 + i) = rD;
        if (INTPOW || p != 0.0) {                                                       // inactive systems contribute nothing to S
            const double w = -si * g;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const double wa = w * pw[a];
#pragma unroll
                for (int b = 0; b <= a; ++b) L.K[a * 6 + b] += wa * pw[b];
            }
        }
        const double srD = si * rD;
#pragma unroll
        for (int c = 0; c < 6; ++c) yr[c] += srD * pw[c];
        if (HASQ && (INTPOW || p != 0.0)) {
            const double spD = si * (k.b_h * p * invD);
#pragma unroll
            for (int c = 0; c < 6; ++c) L.yp[c] += spD * pw[c];
        }
    }
    L.mask = mask;
    L.flip = flip;
    *nrm2_out = nrm2;
    *bad_out = bad;
}

// one pass of the `solve` loop body (nonlinear_solver.jl:53-60): evaluate r and J at x, test, update x.
// returns true when the point is finished (converged, out of iterations, or needs the general path).
// KEEPK: leave L.K = M intact for crystal_finalize (one-kernel path / host twin); otherwise M is factored in place.
template <bool HASQ, bool INTPOW, bool KEEPK, class Scr>
MM_HD bool crystal_iterate(const CrystalK& k, Scr& scr, CrystalLane& L) {
    double yr[6], nrm2;
    bool bad;
    crystal_evaluate<HASQ, INTPOW>(k, scr, L, yr, &nrm2, &bad);
    L.iters = L.it - 1;
    if (L.it > k.max_iter) return true;                                                  // nonlinear_solver.jl:61 (returns nothing)
    if (!(nrm2 == nrm2)) return true;                                                    // NaN residual (NaN input): can never pass :56 -> not converged
    if (bad) { L.bad = true; return true; }                                              // D not safely invertible -> general path
    // norm(r) < tol  (nonlinear_solver.jl:56); the square root is only taken inside a rounding-width band around tol²
    if (nrm2 < k.tol2_lo || (nrm2 <= k.tol2_hi && sqrt(nrm2) < k.tol)) { L.converged = true; return true; }
    // x -= J \ r                                                                        nonlinear_solver.jl:59
    double Mc[KEEPK ? 36 : 1];
    double* M = KEEPK ? Mc : L.K;
    if (KEEPK) {
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) Mc[KEEPK ? (a * 6 + b) : 0] = L.K[a * 6 + b];
    }
    if (!ldl6_factor(M)) { L.bad = true; return true; }
    ldl6_solve(M, yr);
    double fac = 0.0;
    double ypl[6];
    if (HASQ) {
#pragma unroll
        for (int c = 0; c < 6; ++c) ypl[c] = L.yp[c];
        ldl6_solve(M, ypl);
        double su = 0.0, sv = 0.0;
#pragma unroll kCrystalUnrollSmall
        for (int i = 0; i < 12; ++i) {
            const double g = scr(24 + i);
            double du = 0.0, dv = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) { du += k.Pw[i][c] * yr[c]; dv += k.Pw[i][c] * ypl[c]; }
            su += scr(36 + i) + g * du;
            sv += g * dv - k.b_h * fabs(g);
        }
        fac = su * fast_rcp(1.0 - sv);
    }
#pragma unroll kCrystalUnrollSmall
    for (int i = 0; i < 12; ++i) {
        const double g = scr(24 + i);
        double d = scr(36 + i);                       // (D⁻¹r)_i ; the Woodbury correction vanishes for c_i = 0
        double du = 0.0, dv = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) { du += k.Pw[i][c] * yr[c]; if (HASQ) dv += k.Pw[i][c] * ypl[c]; }
        d += g * du;
        if (HASQ) d += (g * dv - k.b_h * fabs(g)) * fac;
        scr(XS + i) -= d;
    }
    L.iters = L.it;
    L.it += 1;
    return false;
}

// ---- SPARSE pass (integer m) ------------------------------------------------------------------------------------------------
// On the benchmark workload only ~1.8 of the 12 systems are active at the solution (step B; 0.9 in step A), yet the branch-free pass
// above runs the full chain — two reciprocals, the power, the rank-1 update of M, ~68 FP64 instructions — for all 12, because with the
// system index uniform across the warp some lane is always active.  Here the system index is PER LANE:
//   1. σ from the systems with x_i != 0 only (a term with x_i = 0 is an exact zero);
//   2. a uniform yield test over the 12 systems (τ_i, Φ_i with α = αₙ, κ = κₙ [+ b_h Σx]: exactly what the full formulas give for
//      x_i = 0, since fma(a_h, 0, κₙ) = κₙ and rcp_nc(1) = 1) -> candidate set = {x_i != 0} ∪ {Φ_i > 0 or NaN};
//   3. the full chain for the candidates only, each lane walking ITS OWN bit list (trip count = the warp's longest list, ~5 instead of 12),
//      the Pw / Eᵉ:MS rows fetched from a shared-memory copy of the tables laid out [component][system] (lane-divergent rows without
//      bank conflicts; the constant bank would serialise them);
//   4. the update for the candidates only — every other system has R_i = −t*·0, g_i = 0, so x_i stays 0.
// Everything skipped is an exact zero in the dense pass, so the iterates are the same numbers up to the flush of 1e-16-level noise
// described at the update (tests: masks, counts and values against the oracle, exactly as for the dense pass).
// tab: [0,72) Pw transposed (tab[c*12 + i]), [72,144) Eᵉ:MS transposed, [144,288) G[i*12 + j] = Pw_i·Eᵉ·Pw_j, symmetrised bit for bit
// (G[i][j] = Gm[min][max]) — the active-set pass below gathers it pair by pair.
constexpr int MM_CRYSTAL_TAB = 288, GTAB = 144;
MM_HD void crystal_fill_tab(const CrystalK& k, double* tab, int e) {      // e in [0,288): one entry (kernel threads fill it cooperatively)
    if (e >= GTAB) { const int i = (e - GTAB) / 12, j = (e - GTAB) % 12; tab[e] = (i <= j) ? k.Gm[i][j] : k.Gm[j][i]; return; }
    const int c = (e % 72) / 12, i = e % 12;
    tab[e] = e < 72 ? k.Pw[i][c] : k.EM[i][c];
}

#if defined(MM_CRYSTAL_STATS) && !defined(__CUDA_ARCH__)
void mm_crystal_stats_pass(unsigned nz, unsigned cand);       // host-twin instrumentation (tests/hosttwin)
#endif
template <bool HASQ, bool KEEPK, bool INTPOW, class Scr> MM_HD bool crystal_sparse_tail(const CrystalK& k, const double* tab, Scr& scr, CrystalLane& L, const unsigned cand, const double kq);
template <bool HASQ, bool KEEPK, bool INTPOW = true, class Scr>
MM_HD bool crystal_iterate_sparse(const CrystalK& k, const double* tab, Scr& scr, CrystalLane& L) {
    double sumx = 0.0;
    // 1. σ = σ_trial − Σ_{x_j != 0} x_j s_j Eᵉ:MS_j                                        :131
#pragma unroll
    for (int c = 0; c < 6; ++c) L.sig[c] = scr(SIGT + c);
    for (unsigned m = L.nz; m != 0u; m &= m - 1u) {
        const int j = __ffs_(m);
        const double xj = scr(XS + j);
        const double f = xj * crystal_sign(L, j);
        if (HASQ) sumx += xj;
#pragma unroll
        for (int c = 0; c < 6; ++c) L.sig[c] -= f * tab[72 + c * 12 + j];
    }
    // 2. uniform yield test
    unsigned cand = L.nz;
    const double kq = HASQ ? k.b_h * sumx : 0.0;
#pragma unroll 4
    for (int i = 0; i < 12; ++i) {
        double tau = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) tau += L.sig[c] * k.Pw[i][c];
        double kap = scr(i);
        if (HASQ) kap += kq;
        const double Phi = fabs(tau - scr(12 + i)) - k.tau_y - kap;
        cand |= (!(Phi <= 0.0)) ? (1u << i) : 0u;                                       // Φ > 0, or NaN (handled by the full chain)
    }
    // systems that were candidates at the last evaluation and are not any more: their g / D⁻¹r slots go back to the zeros the dense pass
    // would write (crystal_finalize reads all twelve); first pass: all slots
    for (unsigned m = L.cand & ~cand; m != 0u; m &= m - 1u) { const int i = __ffs_(m); scr(24 + i) = 0.0; scr(36 + i) = 0.0; }
    L.cand = cand;
#if defined(MM_CRYSTAL_STATS) && !defined(__CUDA_ARCH__)
    mm_crystal_stats_pass(L.nz, cand);
#endif
    return crystal_sparse_tail<HASQ, KEEPK, INTPOW>(k, tab, scr, L, cand, kq);
}

// steps 3 and 4 of the sparse pass (6x6 SPD route): the chains of the candidates, M, the test, the update
template <bool HASQ, bool KEEPK, bool INTPOW, class Scr>
MM_HD bool crystal_sparse_tail(const CrystalK& k, const double* tab, Scr& scr, CrystalLane& L, const unsigned cand, const double kq) {
    double yr[6], nrm2 = 0.0;
    // 3. full chain for the candidates — TWO list entries per trip: one chain is a ~30-level dependency ladder (reciprocal -> α -> Φ -> power ->
    //    D -> reciprocal -> rank-1 update) and three resident warps per sub-partition cannot fill the FP64 pipe with one ladder each.  A lane whose
    //    list has an odd length runs its last entry twice; the second copy contributes exact zeros (w = 0, s·D⁻¹r = 0, R² = 0) and rewrites the
    //    same g / D⁻¹r, so every sum sees the same terms in the same order as one entry per trip.
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) L.K[a * 6 + b] = k.Cmp[a][b];
#pragma unroll
    for (int c = 0; c < 6; ++c) { yr[c] = 0.0; L.yp[c] = 0.0; }
    unsigned mask = 0;
    bool bad = false, flip = false;
    for (unsigned m = cand; m != 0u;) {
        int ii[2];
        ii[0] = __ffs_(m); m &= m - 1u;
        const bool two = m != 0u;
        ii[1] = two ? __ffs_(m) : ii[0]; m &= m - 1u;
        double pw[2][6], w[2], srD[2], spD[2], R2[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = ii[u];
            const bool live = (u == 0) | two;
#pragma unroll
            for (int c = 0; c < 6; ++c) pw[u][c] = tab[c * 12 + i];
            const double si = crystal_sign(L, i), xi = scr(XS + i);
            double kap = scr(i) + k.a_h * xi;                                           // :127 with H = a I + b 11ᵀ
            if (HASQ) kap += kq;
            const double iden = rcp_nc(1.0 + k.hk_ai * xi);
            const double al = (scr(12 + i) + k.H_kin * xi * si) * iden;                 // :128
            double tau = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) tau += L.sig[c] * pw[u][c];
            const double red = tau - al;
            const double sgn = sign_of(red);
            const double Phi = fabs(red) - k.tau_y - kap;                                // :110
            const bool act = Phi > 0.0;
            const double xx = act ? Phi * k.isc : Phi * 0.0;
            const double pw1 = crystal_pow_m1<INTPOW>(k, xx, act);
            const double Ri = k.dt * (pw1 * xx) - k.t_star * xi;                        // :111
            const double p = k.dtm_sc * pw1;
            mask |= (act & live) ? (1u << i) : 0u;
            flip = flip | (act & (sgn != si));
            const double dal = (k.H_kin * si - al * k.hk_ai) * iden;
            const double dD = k.t_star + p * (sgn * dal + k.a_h);
            bad = bad | (act & !(dD >= 0.25 * k.t_star));
            const double invD = act ? -rcp_nc(dD) : k.mits;
            const double ci = p * sgn;
            const double rD = Ri * invD;
            const double g = ci * invD;
            scr(24 + i) = g; scr(36 + i) = rD;
            R2[u] = live ? Ri : 0.0;                                                    // squared at the sum below (same fma as one entry per trip)
            w[u] = live ? -si * g : 0.0;
            srD[u] = live ? si * rD : 0.0;
            spD[u] = (HASQ && live) ? si * (k.b_h * p * invD) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            nrm2 += R2[u] * R2[u];
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const double wa = w[u] * pw[u][a];
#pragma unroll
                for (int b = 0; b <= a; ++b) L.K[a * 6 + b] += wa * pw[u][b];
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) yr[c] += srD[u] * pw[u][c];
            if (HASQ) {
#pragma unroll
                for (int c = 0; c < 6; ++c) L.yp[c] += spD[u] * pw[u][c];
            }
        }
    }
    L.mask = mask;
    L.flip = flip;
    L.iters = L.it - 1;
    if (L.it > k.max_iter) return true;                                                  // nonlinear_solver.jl:61
    if (!(nrm2 == nrm2)) return true;                                                    // NaN residual -> not converged
    if (bad) { L.bad = true; return true; }
    if (nrm2 < k.tol2_lo || (nrm2 <= k.tol2_hi && sqrt(nrm2) < k.tol)) { L.converged = true; return true; }
    // x -= J \ r                                                                        nonlinear_solver.jl:59
    double Mc[KEEPK ? 36 : 1];
    double* M = KEEPK ? Mc : L.K;
    if (KEEPK) {
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) Mc[KEEPK ? (a * 6 + b) : 0] = L.K[a * 6 + b];
    }
    if (!ldl6_factor(M)) { L.bad = true; return true; }
    ldl6_solve(M, yr);
    double fac = 0.0;
    double ypl[6];
    if (HASQ) {
#pragma unroll
        for (int c = 0; c < 6; ++c) ypl[c] = L.yp[c];
        ldl6_solve(M, ypl);
        double su = 0.0, sv = 0.0;
        for (unsigned m = cand; m != 0u; m &= m - 1u) {
            const int i = __ffs_(m);
            const double g = scr(24 + i);
            double du = 0.0, dv = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) { du += tab[c * 12 + i] * yr[c]; dv += tab[c * 12 + i] * ypl[c]; }
            su += scr(36 + i) + g * du;
            sv += g * dv - k.b_h * fabs(g);
        }
        fac = su * fast_rcp(1.0 - sv);
    }
    unsigned nz = 0;
    for (unsigned m = cand; m != 0u;) {                                                // two list entries per trip, as above (an odd tail entry is redone from the same x)
        int ii[2];
        ii[0] = __ffs_(m); m &= m - 1u;
        ii[1] = (m != 0u) ? __ffs_(m) : ii[0]; m &= m - 1u;
        double xn[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = ii[u];
            const double g = scr(24 + i);
            double d = scr(36 + i);
            double du = 0.0, dv = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) { du += tab[c * 12 + i] * yr[c]; if (HASQ) dv += tab[c * 12 + i] * ypl[c]; }
            d += g * du;
            if (HASQ) d += (g * dv - k.b_h * fabs(g)) * fac;
            const double xo = scr(XS + i);
            xn[u] = xo - d;
            // An INACTIVE system's exact Newton step is x_i -> 0 (R_i = −t* x_i, J_ii = −t*); in floating point it lands on rounding noise
            // ~1e-16 |x_i| which would keep the system on the candidate list while it decays over the next ~20 passes.  Noise of that size is
            // flushed to the exact zero the step means (the reference's LU leaves a different 1e-16-level noise there; nothing depends on it).
            if (!((mask >> i) & 1u) && fabs(xn[u]) <= 1e-14 * fabs(xo)) xn[u] = 0.0;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            scr(XS + ii[u]) = xn[u];
            if (xn[u] != 0.0) nz |= (1u << ii[u]);
        }
    }
    L.nz = nz;
    L.iters = L.it;
    L.it += 1;
    return false;
}

// ---- ACTIVE-SET pass: the Newton kernel's default ---------------------------------------------------------------------------------
// The sparse pass still pays, per Newton pass, for the 6x6 machinery of the Woodbury split: 27 FP64 instructions per candidate to accumulate
// M, 6 for y_r, 6 for Pw_i·z, and a 6x6 LDLᵀ + solve (≈ 160, a ladder of six dependent reciprocals) — although on configs[3] a pass has
// 3.0 candidates on average, at most 2 in 47 %, 4 in 83 % and 6 in 96 % of all passes (tools/crystal_sched_sim.py).  Here the Newton step is
// solved where it is small: rows of systems off the candidate list are D_ii δ_i = r_i = 0, and on the list {a_1..a_k}
//        Σ_v (δ_uv + w_u G[a_u][a_v]) δ'_v = s_u (D⁻¹r)_u ,   δ'_v = s_v δ_v ,  w_u = −s_u c_u / D_uu ,  G = Pw Eᵉ Pwᵀ   (table)
// — J restricted to the list, rows scaled by D⁻¹ and by the frozen signs (J_ij = D_ii δ_ij − c_i s_j G_ij, §"Register-resident fast path").
// I + W G is a row scaling of the SPD matrix W⁻¹ + G, so elimination without pivoting is as stable as Cholesky (pivots >= 1 while every
// w_u >= 0); a pivot that is not safely positive sends the point to the pivoted general path like the LDLᵀ guard does.  No reciprocal of w
// is taken: a barely active system (w ~ 1e-200) is simply a row e_u.  τ_i = σ:Pw_i is parked by the yield test, so the chains need no table
// rows at all; they leave g_i and (D⁻¹r)_i in the scratch slots of their system.  The solve is done at ONE size per warp — 2, 4 or 6, the
// smallest that holds the longest list of the warp (a warp-uniform choice: no lane ever runs two routes) — in registers with static
// indexing; a lane with a shorter list pads with rows e_u (w = 0, right-hand side 0: exact no-ops in the elimination, so a point's result
// does not depend on its neighbours).  A point with a list longer than 6 in some pass (7 % of the points of configs[3]: they are known from
// the presort and never enter this pass, bar the few whose list grows), or with a candidate whose frozen sign is 0, is done by the sparse
// pass instead: the kernel around this pass carries no 6x6 code at all (3 360 -> 2 300 instructions; the instruction cache holds 2 000).
// Same Newton iterates up to rounding (masks and counts against the oracle: tests).
// `lanes`: the lanes of the warp that execute this call together (device: ballot of the caller's condition).
// Scratch of the active-set kernel (48 slots per lane, so that 16 warps are resident): [0,12) κₙ [12,24) αₙ [24,36) x=μ, then BY LIST POSITION
// [36,42) w = −s g [42,48) s D⁻¹r; σ_trial stays in registers.
// With cross hardening (q != 0) six more slots [48,54) by list position: b_h p / D (the rank-1 column of the Jacobian).
constexpr int MM_CRYSTAL_SCRATCH_ASET = 48, MM_CRYSTAL_SCRATCH_ASET_Q = 54, XSA = 24, GPA = 36, RPA = 42, HPA = 48;
template <int KM, bool HASQ, class Scr>
MM_HD bool crystal_aset_solve(const double* tab, Scr& scr, CrystalLane& L, const unsigned cand, const unsigned mask) {
    int a[KM];
    double w[KM], b[KM], hs[HASQ ? KM : 1], sgv[HASQ ? KM : 1];
    unsigned rest = cand;
#pragma unroll
    for (int u = 0; u < KM; ++u) {
        const bool live = rest != 0u;
        a[u] = live ? __ffs_(rest) : 0;
        rest &= rest - 1u;
        w[u] = live ? scr(GPA + u) : 0.0;                                               // the chains stored w = −s g, s D⁻¹r (and s b_h p / D) by list position
        b[u] = live ? scr(RPA + u) : 0.0;
        if (HASQ) { hs[HASQ ? u : 0] = live ? scr(HPA + u) : 0.0; sgv[HASQ ? u : 0] = live ? crystal_sign(L, a[u]) : 0.0; }     // cross hardening: row factor s_u b_h p_u / D_uu, column factor s_v
    }
    double Nm[KM][KM];
#pragma unroll
    for (int u = 0; u < KM; ++u)
#pragma unroll
        for (int v = u; v < KM; ++v) Nm[u][v] = tab[GTAB + a[u] * 12 + a[v]];           // G is symmetric bit for bit (crystal_fill_tab)
#pragma unroll
    for (int u = 0; u < KM; ++u)
#pragma unroll
        for (int v = 0; v < u; ++v) Nm[u][v] = w[u] * Nm[v][u];                         // the upper triangle still holds G here
#pragma unroll
    for (int u = 0; u < KM; ++u)
#pragma unroll
        for (int v = u; v < KM; ++v) Nm[u][v] = (u == v) ? fma(w[u], Nm[u][v], 1.0) : w[u] * Nm[u][v];
    if (HASQ) {                                                                         // − (s_u h_u) s_v : the cross-hardening column p̂ 1ᵀ on the list (I + W (G + b_h s sᵀ) while no sign flipped)
#pragma unroll
        for (int u = 0; u < KM; ++u)
#pragma unroll
            for (int v = 0; v < KM; ++v) Nm[u][v] -= hs[HASQ ? u : 0] * sgv[HASQ ? v : 0];
    }
    bool ok = true;
#pragma unroll
    for (int q = 0; q < KM; ++q) {
        const double d = Nm[q][q];
        ok = ok & (d > 1e-200) & (d < 1e200);
        const double id = rcp_nc(d);
        Nm[q][q] = id;
#pragma unroll
        for (int i = q + 1; i < KM; ++i) {
            const double l = Nm[i][q] * id;
#pragma unroll
            for (int j = q + 1; j < KM; ++j) Nm[i][j] -= l * Nm[q][j];
            b[i] -= l * b[q];
        }
    }
    if (!ok) { L.bad = true; return true; }
#pragma unroll
    for (int i = KM - 1; i >= 0; --i) {
        double t = b[i];
#pragma unroll
        for (int j = i + 1; j < KM; ++j) t -= Nm[i][j] * b[j];
        b[i] = t * Nm[i][i];
    }
    // x -= δ on the list (δ_u = s_u δ'_u); every other system keeps x_i = 0
    unsigned nz = 0;
    rest = cand;
#pragma unroll
    for (int u = 0; u < KM; ++u) {
        if (rest != 0u) {
            const int i = a[u];
            const double xo = scr(XSA + i);
            double xn = xo - crystal_apply_sign(L, i, b[u]);
            if (!((mask >> i) & 1u) && fabs(xn) <= 1e-14 * fabs(xo)) xn = 0.0;          // see the sparse pass
            scr(XSA + i) = xn;
            if (xn != 0.0) nz |= (1u << i);
        }
        rest &= rest - 1u;
    }
    L.nz = nz;
    L.iters = L.it;
    L.it += 1;
    return false;
}

template <bool HASQ, bool INTPOW = true, class Scr>
MM_HD bool crystal_iterate_aset(const CrystalK& k, const double* tab, Scr& scr, CrystalLane& L, const double* sigt, const unsigned lanes) {
    // 1. σ = σ_trial − Σ_{x_j != 0} x_j s_j Eᵉ:MS_j                                        :131
#pragma unroll
    for (int c = 0; c < 6; ++c) L.sig[c] = sigt[c];
    double sumx = 0.0;
    for (unsigned m = L.nz; m != 0u; m &= m - 1u) {
        const int j = __ffs_(m);
        const double xj = scr(XSA + j);
        const double f = crystal_apply_sign(L, j, xj);                                  // (a zero frozen sign on the list sends the point to the sparse launch below)
        if (HASQ) sumx += xj;
#pragma unroll
        for (int c = 0; c < 6; ++c) L.sig[c] -= f * tab[72 + c * 12 + j];
    }
    const double kq = HASQ ? k.b_h * sumx : 0.0;                                        // κ_i = κₙ_i + a_h x_i + b_h Σ x   (:127, H = a I + b 11ᵀ)
    // 2. uniform yield test
    unsigned cand = L.nz;
#pragma unroll 4
    for (int i = 0; i < 12; ++i) {
        double tau = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) tau += L.sig[c] * k.Pw[i][c];
        double kap = scr(i);
        if (HASQ) kap += kq;
        const double Phi = fabs(tau - scr(12 + i)) - k.tau_y - kap;
        cand |= (!(Phi <= 0.0)) ? (1u << i) : 0u;                                       // Φ > 0, or NaN (handled by the full chain)
    }
    L.cand = cand;
#if defined(MM_CRYSTAL_STATS) && !defined(__CUDA_ARCH__)
    mm_crystal_stats_pass(L.nz, cand);
#endif
    // the route is a property of the POINT (its result must not depend on its neighbours in the warp): more than 6 candidates in any of its
    // passes, or a candidate whose frozen sign is 0 -> the point is handed to the sparse pass (6x6 Woodbury route) and redone from its first
    // pass there; the solve SIZE below is the warp's (padding is an exact no-op)
#if defined(__CUDA_ARCH__)
    const int kc = __popc(cand);
#else
    const int kc = __builtin_popcount(cand);
#endif
    const bool big = (kc > 6) | ((cand & ~(L.spos | L.sneg)) != 0u);
#if defined(__CUDA_ARCH__)
    const unsigned kmax = __reduce_max_sync(lanes, big ? 0u : (unsigned)kc);            // the longest list among the warp's small-route lanes
#else
    const unsigned kmax = (unsigned)kc; (void)lanes;
#endif
    if (big) { L.big = true; return true; }                                             // the whole point is redone by the sparse pass (second launch)
    // 3. the chains, two list entries per trip (see the sparse pass): g_i, (D⁻¹r)_i into the scratch slots of their system
    double nrm2 = 0.0;
    unsigned mask = 0;
    bool bad = false, flip = false;
    int pos = 0;
    for (unsigned m = cand; m != 0u; pos += 2) {
        int ii[2];
        ii[0] = __ffs_(m); m &= m - 1u;
        const bool two = m != 0u;
        ii[1] = two ? __ffs_(m) : ii[0]; m &= m - 1u;
        double gg[2], rr[2], R2[2], hh[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = ii[u];
            const bool live = (u == 0) | two;
            const double xi = scr(XSA + i);                                             // s_i = ±1 on this list: products with it are sign-bit flips
            double kap = scr(i) + k.a_h * xi;                                           // :127 with H = a I + b 11ᵀ
            if (HASQ) kap += kq;
            const double iden = rcp_nc(1.0 + k.hk_ai * xi);
            const double al = (scr(12 + i) + crystal_apply_sign(L, i, k.H_kin * xi)) * iden;    // :128
            double tau = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) tau += L.sig[c] * tab[c * 12 + i];             // the yield test's τ_i, bit for bit
            const double red = tau - al;
            const double sgn = sign_of(red);
            const double Phi = fabs(red) - k.tau_y - kap;                                // :110
            const bool act = Phi > 0.0;
            const double xx = act ? Phi * k.isc : Phi * 0.0;
            const double pw1 = crystal_pow_m1<INTPOW>(k, xx, act);
            const double Ri = k.dt * (pw1 * xx) - k.t_star * xi;                        // :111
            const double p = k.dtm_sc * pw1;
            mask |= (act & live) ? (1u << i) : 0u;
            flip = flip | (act & ((red == 0.0) | ((red < 0.0) != (((L.sneg >> i) & 1u) != 0u))));       // sgn_i != s_i
            const double dal = (crystal_apply_sign(L, i, k.H_kin) - al * k.hk_ai) * iden;
            const double dD = k.t_star + p * (sgn * dal + k.a_h);
            bad = bad | (act & !(dD >= 0.25 * k.t_star));
            const double invD = act ? -rcp_nc(dD) : k.mits;
            rr[u] = crystal_apply_sign(L, i, Ri * invD);                                // s (D⁻¹r)_u
            gg[u] = -crystal_apply_sign(L, i, (p * sgn) * invD);                        // w_u = −s_u g_u
            hh[u] = HASQ ? crystal_apply_sign(L, i, k.b_h * p * invD) : 0.0;
            R2[u] = live ? Ri : 0.0;
        }
        // by list position (an odd tail entry fills a slot nobody reads)
#pragma unroll
        for (int u = 0; u < 2; ++u) { scr(GPA + pos + u) = gg[u]; scr(RPA + pos + u) = rr[u]; if (HASQ) scr(HPA + pos + u) = hh[u]; nrm2 += R2[u] * R2[u]; }
    }
    L.mask = mask;
    L.flip = flip;
    L.iters = L.it - 1;
    if (L.it > k.max_iter) return true;                                                  // nonlinear_solver.jl:61
    if (!(nrm2 == nrm2)) return true;                                                    // NaN residual -> not converged
    if (bad) { L.bad = true; return true; }
    if (nrm2 < k.tol2_lo || (nrm2 <= k.tol2_hi && sqrt(nrm2) < k.tol)) { L.converged = true; return true; }
    // 4. (I + W G) δ' = s ∘ D⁻¹r on the list, elimination without pivoting                 nonlinear_solver.jl:59
#ifndef MM_ASET_SIZES
#define MM_ASET_SIZES 246      // solve sizes compiled in (tuning macro): 246 = {2, 4, 6}; 6, 26, 36, 46 measured within 0.5 % (profiles/r02_crystal_aset_split_kernels.log)
#endif
#if MM_ASET_SIZES == 246 || MM_ASET_SIZES == 26
    if (kmax <= 2u) return crystal_aset_solve<2, HASQ>(tab, scr, L, cand, mask);
#endif
#if MM_ASET_SIZES == 36
    if (kmax <= 3u) return crystal_aset_solve<3, HASQ>(tab, scr, L, cand, mask);
#endif
#if MM_ASET_SIZES == 246 || MM_ASET_SIZES == 46
    if (kmax <= 4u) return crystal_aset_solve<4, HASQ>(tab, scr, L, cand, mask);
#endif
    return crystal_aset_solve<6, HASQ>(tab, scr, L, cand, mask);
}

// ---- PHASED pass (integer m, q = 0): measured slower, kept selectable (crystal_variant 2) ----------------------------------------------------------------
// ncu on the two passes above (profiles/r02_crystal_dense_vs_sparse_ncu.txt): the dense pass executes 87.7e9 thread-instructions per
// 2e6 points, the sparse one 44.9e9 — half the work — but only 16 % fewer WARP instructions, because the per-lane system lists leave
// 15 of 32 lanes active per instruction.  And in both, ~45 % of the issued instructions are not FP64 math: run-time-indexed constant
// loads of the Pw / Eᵉ:MS rows, scratch traffic, sign extraction.  So the pass is split by the kind of work:
//   1. σ = σ_trial − Σ x_j s_j Eᵉ:MS_j and τ_i = σ:Pw_i for all 12 systems — dense, FULLY UNROLLED, every table entry a compile-time
//      constant-bank operand of the DFMA itself (no loads, no address arithmetic, all 32 lanes busy); the yield test with α = αₙ, κ = κₙ
//      gives the candidate set {x_i != 0} ∪ {Φ_i > 0 or NaN} (exact for x_i = 0, see the sparse pass);
//   2. the scalar chain (two reciprocals, the power, D_ii) — the only part that is wasted on inactive systems — for the candidates only,
//      each lane walking its own list (~3 of 12 per pass); it needs no table at all (τ_i was parked in the D⁻¹r slot);
//   3. M = (Eᵉ)⁻¹ + Σ w_i Pw_i Pw_iᵀ as 21 dot products with the host-built table T[ab][i] = Pw_i[a] Pw_i[b], y_r, ‖r‖² — dense, unrolled;
//   4. LDLᵀ, and the update x_i −= (D⁻¹r)_i + g_i Pw_i·z — dense, unrolled, stores only for the candidates.
// Scratch (72 slots): [0,12) κₙ [12,24) αₙ [24,36) g [36,48) τ -> D⁻¹r [48,60) x [60,72) R_i² (+inf = D_ii guard failed); σ_trial in registers.
struct CrystalNK {
    double Pw[12][6], EM[12][6], T[21][12], Cmp[21];
    double lam, G, tau_y, H_kin, a_h, hk_ai, isc, dtm_sc, dt, t_star, mits, tol, tol2_lo, tol2_hi;
    int max_iter, npow;
};
constexpr int MM_CRYSTAL_SCRATCH_PHASED = 72;
constexpr int R2S = 60;
MM_HD constexpr int tri21(int a, int b) { return a * (a + 1) / 2 + b; }        // b <= a
inline void make_crystal_nk(CrystalNK& n, const CrystalK& k) {
    for (int i = 0; i < 12; ++i)
        for (int c = 0; c < 6; ++c) { n.Pw[i][c] = k.Pw[i][c]; n.EM[i][c] = k.EM[i][c]; }
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b <= a; ++b) {
            n.Cmp[tri21(a, b)] = k.Cmp[a][b];
            for (int i = 0; i < 12; ++i) n.T[tri21(a, b)][i] = k.Pw[i][a] * k.Pw[i][b];
        }
    n.lam = k.lam; n.G = k.G; n.tau_y = k.tau_y; n.H_kin = k.H_kin; n.a_h = k.a_h; n.hk_ai = k.hk_ai; n.isc = k.isc; n.dtm_sc = k.dtm_sc; n.dt = k.dt; n.t_star = k.t_star;
    n.mits = k.mits; n.tol = k.tol; n.tol2_lo = k.tol2_lo; n.tol2_hi = k.tol2_hi; n.max_iter = k.max_iter; n.npow = k.m_int - 1;
}

template <bool KEEPK, class Scr>
MM_HD bool crystal_iterate_phased(const CrystalNK& k, Scr& scr, CrystalLane& L, const double* sigt) {
    // ---- 1. σ, τ_i, yield test
    double sig[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) sig[c] = sigt[c];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        const double f = scr(XS + j) * crystal_sign(L, j);
#pragma unroll
        for (int c = 0; c < 6; ++c) sig[c] -= f * k.EM[j][c];
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) L.sig[c] = sig[c];
    unsigned cand = L.nz;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        double tau = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) tau += sig[c] * k.Pw[i][c];
        scr(36 + i) = tau;
        const double Phi = fabs(tau - scr(12 + i)) - k.tau_y - scr(i);
        cand |= (!(Phi <= 0.0)) ? (1u << i) : 0u;
    }
    for (unsigned m = L.cand & ~cand; m != 0u; m &= m - 1u) scr(24 + __ffs_(m)) = 0.0;     // g slots of systems that left the list (finalize reads all twelve)
    L.cand = cand;
    // ---- 2. scalar chains of the candidates
    unsigned mask = 0;
    bool flip = false;
    for (unsigned m = cand; m != 0u; m &= m - 1u) {
        const int i = __ffs_(m);
        const double si = crystal_sign(L, i), xi = scr(XS + i), tau = scr(36 + i);
        const double kap = scr(i) + k.a_h * xi;                                         // :127 (q = 0)
        const double iden = rcp_nc(1.0 + k.hk_ai * xi);
        const double al = (scr(12 + i) + k.H_kin * xi * si) * iden;                     // :128
        const double red = tau - al;
        const double sgn = sign_of(red);
        const double Phi = fabs(red) - k.tau_y - kap;                                    // :110
        const bool act = Phi > 0.0;
        const double xx = act ? Phi * k.isc : Phi * 0.0;
        const double pw1 = (k.npow == 9) ? powc_<9>(xx) : powi6_(xx, k.npow);
        const double Ri = k.dt * (pw1 * xx) - k.t_star * xi;                            // :111
        const double p = k.dtm_sc * pw1;
        mask |= act ? (1u << i) : 0u;
        flip = flip | (act & (sgn != si));
        const double dal = (k.H_kin * si - al * k.hk_ai) * iden;
        const double dD = k.t_star + p * (sgn * dal + k.a_h);
        const bool bad = act & !(dD >= 0.25 * k.t_star);
        const double invD = act ? -rcp_nc(dD) : k.mits;
        scr(24 + i) = (p * sgn) * invD;
        scr(36 + i) = Ri * invD;
        scr(R2S + i) = bad ? HUGE_VAL : Ri * Ri;
    }
    L.mask = mask;
    L.flip = flip;
    // ---- 3. M, y_r, ‖r‖²
    double K[21], yr[6], nrm2 = 0.0;
#pragma unroll
    for (int e = 0; e < 21; ++e) K[e] = k.Cmp[e];
#pragma unroll
    for (int c = 0; c < 6; ++c) yr[c] = 0.0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const bool in = (cand >> i) & 1u;
        const double si = crystal_sign(L, i);
        const double g = in ? scr(24 + i) : 0.0, rD = in ? scr(36 + i) : 0.0, r2 = in ? scr(R2S + i) : 0.0;
        nrm2 += r2;
        const double w = -si * g, srD = si * rD;
#pragma unroll
        for (int e = 0; e < 21; ++e) K[e] += w * k.T[e][i];
#pragma unroll
        for (int c = 0; c < 6; ++c) yr[c] += srD * k.Pw[i][c];
    }
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) L.K[a * 6 + b] = K[tri21(a, b)];
    L.iters = L.it - 1;
    if (L.it > k.max_iter) return true;                                                  // nonlinear_solver.jl:61
    if (!(nrm2 == nrm2)) return true;                                                    // NaN residual -> not converged
    if (nrm2 == HUGE_VAL) { L.bad = true; return true; }                                 // a D_ii guard failed -> general path
    if (nrm2 < k.tol2_lo || (nrm2 <= k.tol2_hi && sqrt(nrm2) < k.tol)) { L.converged = true; return true; }
    // ---- 4. x -= J \ r                                                                nonlinear_solver.jl:59
    double Mc[KEEPK ? 36 : 1];
    double* M = KEEPK ? Mc : L.K;
    if (KEEPK) {
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) Mc[KEEPK ? (a * 6 + b) : 0] = L.K[a * 6 + b];
    }
    if (!ldl6_factor(M)) { L.bad = true; return true; }
    ldl6_solve(M, yr);
    unsigned nz = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        if ((cand >> i) & 1u) {
            const double g = scr(24 + i), xo = scr(XS + i);
            double du = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) du += k.Pw[i][c] * yr[c];
            const double d = scr(36 + i) + g * du;
            double xn = xo - d;
            if (g == 0.0 && fabs(xn) <= 1e-14 * fabs(xo)) xn = 0.0;                      // inactive system: the exact step is x_i -> 0 (see the sparse pass)
            scr(XS + i) = xn;
            if (xn != 0.0) nz |= (1u << i);
        }
    }
    L.nz = nz;
    L.iters = L.it;
    L.it += 1;
    return false;
}

// outputs of a finished point (:77-90).  Returns the mask (bit 15 = not converged) or *fallback = true.
template <bool HASQ, class Acc, class Scr>
MM_HD unsigned crystal_finalize(const CrystalK& k, Acc& acc, Scr& scr, CrystalLane& L, bool* fallback) {
    if (L.bad) { *fallback = true; return 0; }
    unsigned mask = L.mask;
    if (!L.converged) mask |= 0x8000u;
    // the tangent's SPD factorisation first: if it fails the point goes to the general path and NOTHING may have been written yet
    // (state_out may alias state_in, and the general path restarts from the old state)
    // (factored for every point that writes a tangent, elastic ones included — theirs is unused: a factorisation only the plastic lanes of a warp
    // execute, with an early return inside, left the two halves of the warp un-reconverged to the end of the kernel in one build: 3.75 instead of
    // 2.49 ms per 1e7 points for the same thread-instruction count)
    const bool want_tan = acc.template has_out<1>() && ((mask & 0x0FFFu) != 0) && L.converged;
    bool factored = true;
    if (acc.template has_out<1>()) factored = ldl6_factor(L.K);
    if (want_tan && !factored) { *fallback = true; return 0; }
#pragma unroll
    for (int c = 0; c < 6; ++c) { acc.template out<0>(c, L.sig[c]); acc.template out<2>(c, L.sig[c]); }
    double sumx = 0.0;
    if (HASQ) {
#pragma unroll 1
        for (int j = 0; j < 12; ++j) sumx += scr(XS + j);
    }
#pragma unroll 1
    for (int i = 0; i < 12; ++i) {
        const double xi = scr(XS + i);
        double kap = scr(i) + k.a_h * xi;
        if (HASQ) kap += k.b_h * sumx;
        const double al = (scr(12 + i) + k.H_kin * xi * crystal_sign(L, i)) * fast_rcp(1.0 + k.hk_ai * xi);
        acc.template out<2>(6 + i, kap); acc.template out<2>(18 + i, al); acc.template out<2>(30 + i, xi);
    }
    if (acc.template has_out<1>()) {
        const bool want = ((mask & 0x0FFFu) != 0) && L.converged;
        if (!HASQ) {
            // Closed form (q = 0):  dσdε = Eᵉ + EMᵀ diag(sgn) J⁻¹ diag(c) EM  with  EM = Pw Eᵉ,  J = D − diag(c) Pw Eᵉ Pwᵀ diag(s):
            //   = Eᵉ (I − S_g M⁻¹),  S_g = Σ (p_i/dD_i) Pw_i Pw_iᵀ,  M = (Eᵉ)⁻¹ + S   — and S_g = S unless an active system
            // flipped sign, in which case  dσdε = M⁻¹  (compliance + plastic compliance, inverted).  Six register-resident
            // 6x6 solves replace the six 12-vector Woodbury applications (same algebra; checked against the oracle to 1e-13).
            double Sg[36];
            const bool flip = want && L.flip;
            if (flip) {
#pragma unroll
                for (int a = 0; a < 36; ++a) Sg[a] = 0.0;
#pragma unroll 1
                for (int i = 0; i < 12; ++i) {
                    const double w = fabs(scr(24 + i));                         // p_i / dD_i = |c_i / D_ii|
                    if (w != 0.0) {
#pragma unroll
                        for (int a = 0; a < 6; ++a) {
                            const double wa = w * k.Pw[i][a];
#pragma unroll
                            for (int b = 0; b <= a; ++b) Sg[a * 6 + b] += wa * k.Pw[i][b];
                        }
                    }
                }
            }
            const double G2 = 2.0 * k.G;
#pragma unroll
            for (int Jc = 0; Jc < 6; ++Jc) {
                double col[6];
                {
#pragma unroll
                    for (int Ic = 0; Ic < 6; ++Ic) col[Ic] = (Ic == Jc) ? 1.0 : 0.0;
                    ldl6_solve(L.K, col);                                       // column Jc of M⁻¹ (every lane; an elastic point's column is replaced below)
                    if (flip) {
                        double t[6];                                            // t = S_g m
#pragma unroll
                        for (int a = 0; a < 6; ++a) {
                            double acc_t = 0.0;
#pragma unroll
                            for (int b = 0; b < 6; ++b) acc_t += ((b <= a) ? Sg[a * 6 + b] : Sg[b * 6 + a]) * col[b];
                            t[a] = acc_t;
                        }
                        const double ltr = k.lam * (t[0] + t[3] + t[5]);        // Eᵉ t on raw storage components
#pragma unroll
                        for (int Ic = 0; Ic < 6; ++Ic) {
                            const double et = is_diag_slot(Ic) ? (G2 * t[Ic] + ltr) : (k.G * t[Ic]);
                            col[Ic] = iso_entry(Ic, Jc, k.lam, k.G) - et;
                        }
                    }
                }
#pragma unroll
                for (int Ic = 0; Ic < 6; ++Ic) acc.template out<1>(Ic + 6 * Jc, want ? col[Ic] : iso_entry(Ic, Jc, k.lam, k.G));
            }
            return mask;
        }
        double svq = 0.0;
        if (want) {
            if (HASQ) {            // v = (D − ABᵀ)⁻¹ p̂, kept in scratch [36,48) is taken by the columns -> recomputed per use via yp
                ldl6_solve(L.K, L.yp);
#pragma unroll 1
                for (int i = 0; i < 12; ++i) {
                    const double g = scr(24 + i);
                    double dv = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) dv += k.Pw[i][c] * L.yp[c];
                    svq += g * dv - k.b_h * fabs(g);
                }
            }
        }
        const double ifq = HASQ ? fast_rcp(1.0 - svq) : 0.0;
#pragma unroll 1
        for (int Jc = 0; Jc < 6; ++Jc) {                                                 // one tangent column at a time
            double col[6];
#pragma unroll
            for (int Ic = 0; Ic < 6; ++Ic) col[Ic] = iso_entry(Ic, Jc, k.lam, k.G);
            if (want) {
                // right-hand side ∂R/∂ε column: c_j EM[j][Jc];  Z = J⁻¹ rhs
                double z[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
                for (int j = 0; j < 12; ++j) {
                    const double rD = scr(24 + j) * k.EM[j][Jc];
                    scr(36 + j) = rD;
                    const double srD = crystal_sign(L, j) * rD;
#pragma unroll
                    for (int c = 0; c < 6; ++c) z[c] += srD * k.Pw[j][c];
                }
                ldl6_solve(L.K, z);
                double sz = 0.0;
#pragma unroll 1
                for (int i = 0; i < 12; ++i) {
                    double du = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) du += k.Pw[i][c] * z[c];
                    const double Zi = scr(36 + i) + scr(24 + i) * du;
                    scr(36 + i) = Zi;
                    sz += Zi;
                }
                const double qf = HASQ ? sz * ifq : 0.0;
#pragma unroll 1
                for (int i = 0; i < 12; ++i) {
                    const double g = scr(24 + i);
                    double Zi = scr(36 + i);
                    if (HASQ) {
                        double dv = 0.0;
#pragma unroll
                        for (int c = 0; c < 6; ++c) dv += k.Pw[i][c] * L.yp[c];
                        Zi += (g * dv - k.b_h * fabs(g)) * qf;
                    }
                    const double sgnZ = (g < 0.0) ? Zi : ((g > 0.0) ? -Zi : 0.0);        // sgn_i Z_i (sgn_i = sign c_i = −sign g_i on active systems)
#pragma unroll
                    for (int Ic = 0; Ic < 6; ++Ic) col[Ic] += k.EM[i][Ic] * sgnZ;
                }
            }
#pragma unroll
            for (int Ic = 0; Ic < 6; ++Ic) acc.template out<1>(Ic + 6 * Jc, col[Ic]);
        }
    }
    return mask;
}

// one point start to finish through the SPARSE pass (host twin of crystal_variant 1 and of the heavy-point launch)
template <bool HASQ, class Acc, class Scr>
MM_HD unsigned crystal_point_sparse(const CrystalK& k, Acc& acc, Scr& scr, int* iters_out, bool* fallback) {
    double tab[MM_CRYSTAL_TAB];
    for (int e = 0; e < MM_CRYSTAL_TAB; ++e) crystal_fill_tab(k, tab, e);
    CrystalLane L;
    crystal_init(k, acc, scr, L);
    if (k.m_int) { while (!crystal_iterate_sparse<HASQ, true, true>(k, tab, scr, L)) {} }
    else         { while (!crystal_iterate_sparse<HASQ, true, false>(k, tab, scr, L)) {} }
    const unsigned mask = crystal_finalize<HASQ>(k, acc, scr, L, fallback);
    *iters_out = *fallback ? 0 : L.iters;
    return mask;
}

// one point start to finish through the ACTIVE-SET pass (host twin of the Newton kernel's default path): the Newton
// iteration as the kernel runs it, then — like crystal_finish_kernel — one dense evaluation at the converged x for the outputs
template <bool HASQ, class Acc, class Scr>
MM_HD unsigned crystal_point_aset(const CrystalK& k, Acc& acc, Scr& scr, int* iters_out, bool* fallback) {
    double tab[MM_CRYSTAL_TAB];
    for (int e = 0; e < MM_CRYSTAL_TAB; ++e) crystal_fill_tab(k, tab, e);
    CrystalLane L;
    double sigt[6];
    crystal_init_t<XSA, -1>(k, acc, scr, L, sigt);
    if (k.m_int) { while (!crystal_iterate_aset<HASQ, true>(k, tab, scr, L, sigt, 0xffffffffu)) {} }
    else         { while (!crystal_iterate_aset<HASQ, false>(k, tab, scr, L, sigt, 0xffffffffu)) {} }
    if (L.big) return crystal_point_sparse<HASQ>(k, acc, scr, iters_out, fallback);          // redone from the start, as the second launch does
    if (L.bad) { *fallback = true; *iters_out = 0; return 0; }
    {   // the dense evaluation below uses the standard layout: move x, restore σ_trial
        double xs[12];
        for (int i = 0; i < 12; ++i) xs[i] = scr(XSA + i);
        for (int i = 0; i < 12; ++i) scr(XS + i) = xs[i];
        for (int c = 0; c < 6; ++c) scr(SIGT + c) = sigt[c];
    }
    const bool converged = L.converged;
    const int iters = L.iters;
    double yr[6], nrm2;
    bool bad;
    if (k.m_int) crystal_evaluate<HASQ, true>(k, scr, L, yr, &nrm2, &bad);
    else         crystal_evaluate<HASQ, false>(k, scr, L, yr, &nrm2, &bad);
    if (bad) { *fallback = true; *iters_out = 0; return 0; }
    L.converged = converged;
    const unsigned mask = crystal_finalize<HASQ>(k, acc, scr, L, fallback);
    *iters_out = *fallback ? 0 : iters;
    return mask;
}

// one point start to finish through the PHASED pass (host twin of crystal_variant 2; integer m, q = 0)
template <class Acc, class Scr>
MM_HD unsigned crystal_point_phased(const CrystalK& k, Acc& acc, Scr& scr, int* iters_out, bool* fallback) {
    CrystalNK nk;
    make_crystal_nk(nk, k);
    CrystalLane L;
    crystal_init(k, acc, scr, L);
    double sigt[6];
    for (int c = 0; c < 6; ++c) sigt[c] = scr(SIGT + c);
    while (!crystal_iterate_phased<true>(nk, scr, L, sigt)) {}
    for (int c = 0; c < 6; ++c) scr(SIGT + c) = sigt[c];          // the R² slots overlap σ_trial's; crystal_finalize does not read either
    const unsigned mask = crystal_finalize<false>(k, acc, scr, L, fallback);
    *iters_out = *fallback ? 0 : L.iters;
    return mask;
}

// one point start to finish (host twin, and the non-persistent wrapper)
template <bool HASQ, bool INTPOW, class Acc, class Scr>
MM_HD unsigned crystal_point_fast_t(const CrystalK& k, Acc& acc, Scr& scr, int* iters_out, bool* fallback) {
    CrystalLane L;
    crystal_init(k, acc, scr, L);
    while (!crystal_iterate<HASQ, INTPOW, true>(k, scr, L)) {}
    const unsigned mask = crystal_finalize<HASQ>(k, acc, scr, L, fallback);
    *iters_out = *fallback ? 0 : L.iters;
    return mask;
}
template <bool HASQ, class Acc, class Scr>
MM_HD unsigned crystal_point_fast(const CrystalK& k, Acc& acc, Scr& scr, int* iters_out, bool* fallback) {
    return k.m_int ? crystal_point_fast_t<HASQ, true>(k, acc, scr, iters_out, fallback) : crystal_point_fast_t<HASQ, false>(k, acc, scr, iters_out, fallback);
}

// Second half of the two-kernel path: the Newton kernel left the converged μ in the state output; re-evaluate there
// (same code as the last Newton pass, so σ, M, masks are reproduced) and write all outputs — one thread per point,
// no divergence between neighbouring points.  `acc` must offer rd_out<2>(c) (read back from the state output).
template <bool HASQ, bool INTPOW, class Acc, class Scr>
MM_HD unsigned crystal_finish_from_x(const CrystalK& k, Acc& acc, Scr& scr, bool converged, bool* fallback) {
    CrystalLane L;
    crystal_init(k, acc, scr, L);
#pragma unroll
    for (int i = 0; i < 12; ++i) scr(XS + i) = acc.template rd_out<2>(30 + i);
    double yr[6], nrm2;
    bool bad;
    crystal_evaluate<HASQ, INTPOW>(k, scr, L, yr, &nrm2, &bad);
    if (bad) { *fallback = true; return 0; }
    L.converged = converged;
    return crystal_finalize<HASQ>(k, acc, scr, L, fallback);
}

}  // namespace mm
