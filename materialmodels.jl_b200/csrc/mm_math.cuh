// Per-point constitutive updates (FP64), written once as accessor-templated inline functions.
// The CUDA kernels (mm_kernels.cuh) instantiate them with coalesced-SoA or smem-staged-AoS
// accessors; tests/hosttwin instantiates the same source with a host accessor to check the
// algebra on a CPU box (a debugging aid of the test-suite only — the shipped library exports no
// CPU compute path).
//
// Storage orders follow Tensors.jl (see include/mmb200.h): symmetric 2nd order = (11,21,31,22,32,33),
// full 2nd order = column-major 9, minor-symmetric 4th order = I + 6 J.
//
// Accessor concept:   double acc.template in<A>(c)            read component c of input array A
//                     void   acc.template out<A>(c, v)        write component c of output array A
//                     bool   acc.template has_out<A>()        output array A requested (non-NULL)
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MM_HD __host__ __device__ __forceinline__
#else
#define MM_HD inline
#endif

namespace mm {

// ---- small symmetric-tensor helpers on double[6] = (11,21,31,22,32,33) ---------------------------
MM_HD double ddot6(const double* a, const double* b) {
    return a[0] * b[0] + a[3] * b[3] + a[5] * b[5] + 2.0 * (a[1] * b[1] + a[2] * b[2] + a[4] * b[4]);
}
MM_HD double tr6(const double* a) { return a[0] + a[3] + a[5]; }

// Reciprocal / reciprocal square root for well-scaled positive normal arguments, <= ~1 ulp, WITHOUT the IEEE
// slow-path scaffolding nvcc emits around `1.0/x` and `sqrt` (sub-normal / inf handling: a BSSY/CALL/BSYNC region
// plus ~10 integer instructions each; ncu showed 56 % of the J2 kernel's issued instructions were not FP64 math).
// MUFU.RCP64H / MUFU.RSQ64H seed (~2^-20) + Newton steps in DFMA.  Host build: plain division / sqrt.
MM_HD double fast_rcp(double x) {
#if defined(__CUDA_ARCH__)
    if (!(fabs(x) >= 1e-290 && fabs(x) <= 1e290)) return 1.0 / x;      // sub-normal / huge / NaN: IEEE path (never on sane data)
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
// returns 1/sqrt(x); *root = sqrt(x)
MM_HD double fast_rsqrt(double x, double* root) {
#if defined(__CUDA_ARCH__)
    if (!(x >= 1e-290 && x <= 1e290)) { const double r = sqrt(x); *root = r; return 1.0 / r; }   // 0, sub-normal, huge, NaN
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);        // ½(1 − x y²)
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    *root = x * y;
    return y;
#else
    const double r = sqrt(x);
    *root = r;
    return 1.0 / r;
#endif
}

MM_HD void dev6(const double* a, double* d) {
    const double m = tr6(a) * (1.0 / 3.0);
    d[0] = a[0] - m; d[1] = a[1]; d[2] = a[2]; d[3] = a[3] - m; d[4] = a[4]; d[5] = a[5] - m;
}
MM_HD bool is_diag_slot(int c) { return c == 0 || c == 3 || c == 5; }

// isotropic stiffness entry Eᵉ[I,J] (raw components) — reference src/LinearElastic.jl:27-35
MM_HD double iso_entry(int I, int J, double lam, double G) {
    const bool dI = is_diag_slot(I), dJ = is_diag_slot(J);
    if (dI && dJ) return (I == J) ? (lam + 2.0 * G) : lam;
    return (I == J) ? G : 0.0;
}
// deviatoric symmetric projector entry I_dev[I,J] (raw components)
MM_HD double idev_entry(int I, int J) {
    const bool dI = is_diag_slot(I), dJ = is_diag_slot(J);
    if (dI && dJ) return (I == J) ? (2.0 / 3.0) : (-1.0 / 3.0);
    return (I == J) ? 0.5 : 0.0;
}

// =================================================================================================
// LinearElastic — reference src/LinearElastic.jl:57-60:  σ = Eᵉ ⊡ ε, tangent = Eᵉ
// in<0> = ε(6);  out<0> = σ(6), out<1> = tangent(36, optional)
// =================================================================================================
struct ElasticK { double lam, G; };
MM_HD ElasticK make_elastic_k(double E, double nu) {
    ElasticK k;
    k.lam = E * nu / ((1 + nu) * (1 - 2 * nu));
    k.G = E / (2 * (1 + nu));
    return k;
}
template <class Acc> MM_HD void elastic_tangent_out(const ElasticK& k, Acc& acc) {
#pragma unroll
    for (int J = 0; J < 6; ++J)
#pragma unroll
        for (int I = 0; I < 6; ++I) acc.template out<1>(I + 6 * J, iso_entry(I, J, k.lam, k.G));
}
template <class Acc> MM_HD void linear_elastic_point(const ElasticK& k, Acc& acc) {
    double e[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) e[c] = acc.template in<0>(c);
    const double ltr = k.lam * tr6(e);
    const double G2 = 2.0 * k.G;
#pragma unroll
    for (int c = 0; c < 6; ++c) acc.template out<0>(c, is_diag_slot(c) ? (G2 * e[c] + ltr) : (G2 * e[c]));
    if (acc.template has_out<1>()) elastic_tangent_out(k, acc);
}

// =================================================================================================
// Plastic — reference src/Plastic.jl:126-172.
// in<0> = ε(6) total strain, in<1> = state [εᵖ(6), κ, α(6), μ];
// out<0> = σ(6), out<1> = ∂σ∂ε(36, optional), out<2> = new state(14)
//
// The reference solves the 14-unknown system R(σ,κ,α,μ) = 0 with NLsolve's plain Newton and a dense
// ForwardDiff Jacobian, then takes the tangent from the top-left 6x6 of inv(J).  Here the SAME Newton
// iterates are produced by a closed-form solve of J dx = -R: with s = dev(σ-α), n = s/|s|, ν̂ = √(3/2) n,
// every block of J is a combination of the identity, the deviatoric projector and n⊗n (Eᵉ:N = 2G N for
// deviatoric N), so the step splits into a volumetric part, the component along n, the deviatoric part
// orthogonal to n and two scalar equations (derivation in DESIGN.md §4.2).  Same algebra, ~150 flops per
// iteration instead of a 14x14 LU, no matrix in registers.
// =================================================================================================
struct PlasticK {
    double lam, G;                  // Lamé constants of Eᵉ
    double sigma_y, H_kappa, H_alpha, kappa_inf, c2;   // c2 = 3/(2 α_∞)
    double ikinf, hk_ikinf;         // 1/κ_∞, H_κ/κ_∞
    double tol; int max_iter;
};
MM_HD PlasticK make_plastic_k(double E, double nu, double sigma_y, double H, double r, double kappa_inf, double alpha_inf, double tol, int max_iter) {
    PlasticK k;
    k.lam = E * nu / ((1 + nu) * (1 - 2 * nu));
    k.G = E / (2 * (1 + nu));
    k.sigma_y = sigma_y;
    k.H_kappa = -r * H;                     // Plastic.jl:31
    k.H_alpha = -2.0 / 3.0 * (1 - r) * H;   // Plastic.jl:32
    k.kappa_inf = kappa_inf;
    k.ikinf = 1.0 / kappa_inf;
    k.hk_ikinf = k.H_kappa / kappa_inf;
    k.c2 = 3.0 / (2 * alpha_inf);
    k.tol = tol; k.max_iter = max_iter;
    return k;
}

// returns flag bits (MM_FLAG_PLASTIC=1, MM_FLAG_FAILED=0x80); *iters = number of Newton updates
//
// Control flow is shaped for SIMT: the yield decision selects who runs the Newton loop, but ALL loads happen
// before it and ALL stores after it, executed by the full warp (elastic lanes carry ξ = 0, u = 0 so the one
// tangent formula Eᵉ − 2Gξ I_dev + u⊗n degenerates to Eᵉ exactly).  A per-branch store phase would issue every
// store twice with complementary half-empty lane masks — partially written 32 B sectors, measured 2.2x slower.
// the three ingredients of the consistent tangent  ∂σ∂ε = Eᵉ − gxi I_dev + u ⊗ n  (elastic point: gxi = 0, u = 0)
struct PlasticTangentParts { double gxi; double u[6], n[6]; };
// Spelled with explicit fma so that the device kernels and the host-side expander of the compact return (mm_expand.cpp, which
// rebuilds the dense tangent from the stored {gxi, u, n}) round identically: both evaluate fma(u_I, n_J, fma(-gxi, I_dev, Eᵉ)).
MM_HD double plastic_tangent_entry_lg(double lam, double G, double gxi, const double* u, const double* n, int I, int J) {
    return fma(u[I], n[J], fma(-gxi, idev_entry(I, J), iso_entry(I, J, lam, G)));
}
MM_HD double plastic_tangent_entry(const PlasticK& k, const PlasticTangentParts& tp, int I, int J) {
    return plastic_tangent_entry_lg(k.lam, k.G, tp.gxi, tp.u, tp.n, I, J);
}
// plastic_point_parts: the update without the 36 tangent stores (the dimension wrappers evaluate only the entries they need per outer pass)
template <class Acc> MM_HD int plastic_point_parts(const PlasticK& k, Acc& acc, int* iters_out, PlasticTangentParts& tp) {
    const double c1 = 1.2247448713915890491;   // sqrt(3/2)
    const double r2 = 1.4142135623730950488;   // sqrt(2) (Mandel scaling of the shear residuals)
    const double G2 = 2.0 * k.G;
    double ep[6], st[6], al_n[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) ep[c] = acc.template in<1>(c);
    const double kap_n = acc.template in<1>(6);
#pragma unroll
    for (int c = 0; c < 6; ++c) al_n[c] = acc.template in<1>(7 + c);
    const double mu_n = acc.template in<1>(13);
    {   // σ_trial = Eᵉ ⊡ (ε − εᵖ)                                                        Plastic.jl:131
        double e[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) e[c] = acc.template in<0>(c) - ep[c];
        const double ltr = k.lam * tr6(e);
#pragma unroll
        for (int c = 0; c < 6; ++c) st[c] = is_diag_slot(c) ? (G2 * e[c] + ltr) : (G2 * e[c]);
    }
    double d6[6], s[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) d6[c] = st[c] - al_n[c];
    dev6(d6, s);
    double rho, inv_rho;
    inv_rho = fast_rsqrt(ddot6(s, s), &rho);
    const double Phi = c1 * rho - k.sigma_y - kap_n;                                   // Plastic.jl:132
    // result registers, pre-set to the elastic answer (σ_trial, Eᵉ, state unchanged)     Plastic.jl:134-135
    double sg[6], al[6], n[6], u[6], kap = kap_n, mu = mu_n, gxi = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) { sg[c] = st[c]; al[c] = al_n[c]; n[c] = 0.0; u[c] = 0.0; }
    int iters = 0, flag = 0;
    if (!(Phi <= 0.0)) {                                                                // Plastic.jl:134 `if Φ <= 0` (a NaN Φ takes the Newton branch, as in the reference)
        // ---- Newton on x = (σ, κ, α, μ), x0 = (σ_trial, κₙ, αₙ, μₙ)                   Plastic.jl:142-148
        flag = 1;
        double w[6], i1psi, theta, i1eta, iden, ikk, gk;
        for (;;) {
            // residuals at x                                                           Plastic.jl:162-172
            if (iters > 0) {
#pragma unroll
                for (int c = 0; c < 6; ++c) d6[c] = sg[c] - al[c];
                dev6(d6, s);
                inv_rho = fast_rsqrt(ddot6(s, s), &rho);
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) n[c] = s[c] * inv_rho;
            dev6(al, w);
            const double kr = kap * k.ikinf - 1.0;
            double Rs[6], Ra[6];
            const double gm = G2 * mu * c1, hm = mu * k.H_alpha;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                Rs[c] = (sg[c] - st[c]) + gm * n[c];
                Ra[c] = (al[c] - al_n[c]) - hm * (k.c2 * w[c] - c1 * n[c]);
            }
            const double Rk = (kap - kap_n) - mu * k.H_kappa * kr;
            const double Rm = c1 * rho - k.sigma_y - kap;
            // NLsolve convergence test: max|F| <= ftol on the MANDEL vector (shear entries carry sqrt(2))
            bool conv = (fabs(Rk) <= k.tol) & (fabs(Rm) <= k.tol);
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const double t = is_diag_slot(c) ? k.tol : k.tol * (1.0 / r2);
                conv = conv & (fabs(Rs[c]) <= t) & (fabs(Ra[c]) <= t);
            }
            // quantities shared by the step and by the tangent (two reciprocals serve four quotients)
            const double beta = c1 * inv_rho;
            const double a1 = 1.0 - hm * k.c2;                 // 1 − ψ
            const double a2 = 1.0 - mu * k.hk_ikinf;           // k_κ
            const double ia = fast_rcp(a1 * a2);
            i1psi = a2 * ia;
            ikk = a1 * ia;
            theta = gm * inv_rho;                              // 2G μ β
            const double phi = hm * beta;
            const double eta1 = 1.0 + (theta - phi * i1psi);   // 1 + η
            gk = -k.H_kappa * kr;
            double gt[6], ha[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                ha[c] = k.H_alpha * (c1 * n[c] - k.c2 * w[c]);
                gt[c] = G2 * c1 * n[c] - ha[c] * i1psi;
            }
            const double n_gt = ddot6(n, gt);
            const double den = c1 * n_gt - gk * ikk;
            const double ib = fast_rcp(den * eta1);
            i1eta = den * ib;
            iden = eta1 * ib;
            if (conv) break;                                              // f_converged
            // out of iterations, or a non-finite residual (NaN/Inf input: NLsolve raises on it; here the point is flagged as failed)
            if (iters >= k.max_iter || !(fabs(Rm) < 1.7e308)) { flag |= 0x80; break; }
            // ---- closed-form Newton step
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
                sg[c] += -Rs[c] - theta * Pd - G2 * c1 * n[c] * dmu;
                al[c] += (-ra[c] - phi * Pd - ha[c] * dmu) * i1psi - trRa * dl;
            }
            kap += dkap;
            mu += dmu;
            ++iters;
        }
        // ---- results at the converged iterate                                        Plastic.jl:150-155
        const double mnu = mu * c1;    // εᵖ = εᵖₙ + μ ν̂,  ν̂ = (3/(2√(3/2)|s|)) s = √(3/2) n
#pragma unroll
        for (int c = 0; c < 6; ++c) ep[c] += mnu * n[c];
        // ∂σ∂ε = (J⁻¹)_σσ ⊡ Eᵉ = Eᵉ − 2Gξ I_dev + u ⊗ n                                  Plastic.jl:152-154
        const double xi = theta * i1eta;
        const double omega = k.H_alpha * k.c2 * i1psi;
        const double cD = c1 * iden;
        const double nw = ddot6(n, w);
#pragma unroll
        for (int c = 0; c < 6; ++c) u[c] = G2 * (xi * n[c] + cD * (xi * omega * (w[c] - nw * n[c]) - G2 * c1 * n[c]));
        gxi = G2 * xi;
    }
    // ---- store phase, whole warp converged
#pragma unroll
    for (int c = 0; c < 6; ++c) acc.template out<0>(c, sg[c]);
#pragma unroll
    for (int c = 0; c < 6; ++c) acc.template out<2>(c, ep[c]);
    acc.template out<2>(6, kap);
#pragma unroll
    for (int c = 0; c < 6; ++c) acc.template out<2>(7 + c, al[c]);
    acc.template out<2>(13, mu);
    tp.gxi = gxi;
#pragma unroll
    for (int c = 0; c < 6; ++c) { tp.u[c] = u[c]; tp.n[c] = n[c]; }
    *iters_out = iters;
    return flag;
}
template <class Acc> MM_HD int plastic_point(const PlasticK& k, Acc& acc, int* iters_out) {
    PlasticTangentParts tp;
    const int flag = plastic_point_parts(k, acc, iters_out, tp);
    if (acc.template has_out<1>()) {
#pragma unroll
        for (int J = 0; J < 6; ++J)
#pragma unroll
            for (int I = 0; I < 6; ++I) acc.template out<1>(I + 6 * J, plastic_tangent_entry(k, tp, I, J));
    }
    return flag;
}

// =================================================================================================
// Hyperelasticity — reference src/FiniteStrain/{neohook.jl:35-43, yeoh.jl:37-54, stvenant.jl:36-48};
// from F via C = tdot(F) = FᵀF (src/traits.jl:35-38, :91-130).
// in<0> = C(6) or F(9);  out<0> = S(6), out<1> = ∂S∂C(36, optional)
// All three models have the form  S = a I + b C⁻¹ + g C,
//   ∂S∂C = p I⊗I + q sym(C⁻¹ ⊗̄ C⁻¹) + r C⁻¹⊗C⁻¹   (StVenant: the same with C⁻¹ := I)
// Yeoh: the reference inverts the 9x9 matrix of otimesu(-C, C) numerically (yeoh.jl:44); analytically
// inv(-C ⊗̄ C) = -C⁻¹ ⊗̄ C⁻¹, which is what is evaluated here.
// =================================================================================================
struct HyperK { double lambda, mu, c2, c3; };
// KIND: 0 = C (RightCauchyGreen), 1 = F (DeformationGradient; C = FᵀF, traits.jl:35-38), 2 = E (GreenLagrange; C = 2E + I, traits.jl:43-46)
// TAN : 0 = (S, ∂S∂C) native, 1 = (S, ∂S∂E = 2 ∂S∂C) traits.jl:61-63, 2 = (Pᵀ = S⋅Fᵀ, ∂Pᵀ∂F) traits.jl:67-77 (needs KIND = 1):
//       ∂Pᵀ∂F_ijkl = S_il δ_jk + 2 Σ_n Σ_q F_jn ∂S∂C_inlq F_kq   (the reference's otimesl/otimesu expression, minor symmetry used)
//       out<0> = Pᵀ(9, column-major), out<1> = ∂Pᵀ∂F(81, i + 3j + 9k + 27l)
template <int MODEL, int KIND, int TAN, class Acc> MM_HD void hyper_point_ex(const HyperK& k, Acc& acc) {
    double C[6];
    double F[9];
    if (KIND == 1) {   // F (column-major 9) -> C = FᵀF
#pragma unroll
        for (int c = 0; c < 9; ++c) F[c] = acc.template in<0>(c);
        C[0] = F[0] * F[0] + F[1] * F[1] + F[2] * F[2];
        C[1] = F[3] * F[0] + F[4] * F[1] + F[5] * F[2];
        C[2] = F[6] * F[0] + F[7] * F[1] + F[8] * F[2];
        C[3] = F[3] * F[3] + F[4] * F[4] + F[5] * F[5];
        C[4] = F[6] * F[3] + F[7] * F[4] + F[8] * F[5];
        C[5] = F[6] * F[6] + F[7] * F[7] + F[8] * F[8];
    } else if (KIND == 2) {
#pragma unroll
        for (int c = 0; c < 6; ++c) C[c] = is_diag_slot(c) ? (2.0 * acc.template in<0>(c) + 1.0) : (2.0 * acc.template in<0>(c));
    } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) C[c] = acc.template in<0>(c);
    }
    double A[6];          // C⁻¹ (or I for StVenant)
    double a, b, g, p, q, r;
    if (MODEL == 2) {     // StVenant: S = λ/2 (tr C − 3) I + μ (C − I);  ∂S∂C = ½λ I⊗I + μ I_sym
        A[0] = 1.0; A[1] = 0.0; A[2] = 0.0; A[3] = 1.0; A[4] = 0.0; A[5] = 1.0;
        a = k.lambda / 2 * (tr6(C) - 3.0) - k.mu; b = 0.0; g = k.mu;
        p = 0.5 * k.lambda; q = k.mu; r = 0.0;
    } else {
        // C = [c0 c1 c2; c1 c3 c4; c2 c4 c5];  adjugate / det
        const double m00 = C[3] * C[5] - C[4] * C[4];
        const double m01 = C[2] * C[4] - C[1] * C[5];
        const double m02 = C[1] * C[4] - C[2] * C[3];
        const double det = C[0] * m00 + C[1] * m01 + C[2] * m02;
        const double idet = fast_rcp(det);
        A[0] = m00 * idet; A[1] = m01 * idet; A[2] = m02 * idet;
        A[3] = (C[0] * C[5] - C[2] * C[2]) * idet;
        A[4] = (C[1] * C[2] - C[0] * C[4]) * idet;
        A[5] = (C[0] * C[3] - C[1] * C[1]) * idet;
        const double lnJ = 0.5 * log(det);          // ln J = ln √det C, without the square root
        g = 0.0;
        if (MODEL == 0) {   // NeoHook: S = μ (I − C⁻¹) + λ ln J C⁻¹
            a = k.mu; b = k.lambda * lnJ - k.mu;
            p = 0.0; q = k.mu - k.lambda * lnJ; r = 0.5 * k.lambda;
        } else {            // Yeoh
            const double i3 = tr6(C) - 3.0;
            a = k.mu + 4.0 * k.c2 * i3 + 6.0 * k.c3 * (i3 * i3); b = k.lambda * lnJ - k.mu;
            p = 4.0 * k.c2 + 12.0 * k.c3 * i3; q = k.mu - k.lambda * lnJ; r = 0.5 * k.lambda;
        }
    }
    double S[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        S[c] = b * A[c] + g * C[c];
        if (is_diag_slot(c)) S[c] += a;
    }
    const int SI[6] = {0, 1, 2, 1, 2, 2}, SJ[6] = {0, 0, 0, 1, 1, 2};
    const int SIDX[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    if (TAN != 2) {
        const double tf = (TAN == 1) ? 2.0 : 1.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) acc.template out<0>(c, S[c]);
        if (acc.template has_out<1>()) {
#pragma unroll
            for (int J = 0; J < 6; ++J)
#pragma unroll
                for (int I = 0; I < 6; ++I) {
                    const int i = SI[I], j = SJ[I], kk = SI[J], l = SJ[J];
                    const double sym_u = 0.5 * (A[SIDX[i][kk]] * A[SIDX[j][l]] + A[SIDX[i][l]] * A[SIDX[j][kk]]);
                    double v = q * sym_u + r * (A[I] * A[J]);
                    if (i == j && kk == l) v += p;
                    acc.template out<1>(I + 6 * J, tf * v);
                }
        }
    } else {
        // Pᵀ_ij = Σ_n S_in F_jn
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 3; ++i)
                acc.template out<0>(i + 3 * j, S[SIDX[i][0]] * F[j] + S[SIDX[i][1]] * F[j + 3] + S[SIDX[i][2]] * F[j + 6]);
        if (acc.template has_out<1>()) {
            double D[36];       // ∂S∂C (minor symmetric)
#pragma unroll
            for (int J = 0; J < 6; ++J)
#pragma unroll
                for (int I = 0; I < 6; ++I) {
                    const int i = SI[I], j = SJ[I], kk = SI[J], l = SJ[J];
                    const double sym_u = 0.5 * (A[SIDX[i][kk]] * A[SIDX[j][l]] + A[SIDX[i][l]] * A[SIDX[j][kk]]);
                    double v = q * sym_u + r * (A[I] * A[J]);
                    if (i == j && kk == l) v += p;
                    D[I + 6 * J] = v;
                }
#pragma unroll
            for (int l = 0; l < 3; ++l)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    double G[3][3];     // G[n][k] = Σ_q D_(in)(lq) F_kq
#pragma unroll
                    for (int n = 0; n < 3; ++n)
#pragma unroll
                        for (int kk = 0; kk < 3; ++kk)
                            G[n][kk] = D[SIDX[i][n] + 6 * SIDX[l][0]] * F[kk] + D[SIDX[i][n] + 6 * SIDX[l][1]] * F[kk + 3] + D[SIDX[i][n] + 6 * SIDX[l][2]] * F[kk + 6];
#pragma unroll
                    for (int kk = 0; kk < 3; ++kk)
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            double v = 2.0 * (F[j] * G[0][kk] + F[j + 3] * G[1][kk] + F[j + 6] * G[2][kk]);
                            if (j == kk) v += S[SIDX[i][l]];
                            acc.template out<1>(i + 3 * j + 9 * kk + 27 * l, v);
                        }
                }
        }
    }
}
template <int MODEL, int KIND, class Acc> MM_HD void hyper_point(const HyperK& k, Acc& acc) { hyper_point_ex<MODEL, KIND, 0>(k, acc); }

// =================================================================================================
// XuNeedleman — reference src/CohesiveMaterials/XuNeedleman.jl:75-91.  T = ∂Φ/∂Δ, dTdΔ = ∂²Φ/∂Δ²
// of the potential at :86 (the reference differentiates it with nested duals; closed forms here).
// in<0> = Δ(dim), last component normal;  out<0> = T(dim), out<1> = dTdΔ(dim², column-major, optional)
// =================================================================================================
struct XuK { double Phi_n, delta_n, delta_t, q, r, idn, idt2, A0, B0; };
MM_HD XuK make_xu_k(double Phi_n, double Phi_t, double delta_n, double delta_t, double Delta_n_star) {
    XuK k; k.Phi_n = Phi_n; k.delta_n = delta_n; k.delta_t = delta_t;
    k.q = Phi_t / Phi_n;              // :83
    k.r = Delta_n_star / delta_n;     // :84
    // per-material constants of the potential, hoisted out of the per-point code (FP64 divisions cost ~40 instructions each)
    k.idn = 1.0 / delta_n; k.idt2 = 1.0 / (delta_t * delta_t);
    k.A0 = (1 - k.q) / (k.r - 1); k.B0 = (k.r - k.q) / (k.r - 1);
    return k;
}
template <int DIM, class Acc> MM_HD void xu_point(const XuK& k, Acc& acc) {
    double D[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) D[c] = acc.template in<0>(c);
    const double a = D[DIM - 1] * k.idn;
    double t2 = 0.0;
#pragma unroll
    for (int c = 0; c < DIM - 1; ++c) t2 += D[c] * D[c];
    const double idt2 = k.idt2;
    const double E1 = exp(-a), E2 = exp(-t2 * idt2);
    const double A0 = k.A0, B0 = k.B0;
    const double f = (1 - k.r + a) * A0, gq = k.q + B0 * a;
    const double pe = k.Phi_n * E1;
    // gradient
    const double Tn = pe * k.idn * ((A0 - f) + (gq - B0) * E2);
    const double ct = 2.0 * pe * gq * E2 * idt2;     // T_t = ct Δ_t
#pragma unroll
    for (int c = 0; c < DIM - 1; ++c) acc.template out<0>(c, ct * D[c]);
    acc.template out<0>(DIM - 1, Tn);
    if (acc.template has_out<1>()) {
        const double Hnn = pe * (k.idn * k.idn) * ((f - 2.0 * A0) + (2.0 * B0 - gq) * E2);
        const double cnt = -2.0 * pe * k.idn * (gq - B0) * E2 * idt2;   // H_nt = cnt Δ_t
#pragma unroll
        for (int j = 0; j < DIM - 1; ++j) {
#pragma unroll
            for (int i = 0; i < DIM - 1; ++i) {
                const double v = ct * ((i == j ? 1.0 : 0.0) - 2.0 * D[i] * D[j] * idt2);
                acc.template out<1>(i + DIM * j, v);
            }
            acc.template out<1>((DIM - 1) + DIM * j, cnt * D[j]);
            acc.template out<1>(j + DIM * (DIM - 1), cnt * D[j]);
        }
        acc.template out<1>((DIM - 1) + DIM * (DIM - 1), Hnn);
    }
}

// =================================================================================================
// TransverselyIsotropic — reference src/transverselyisotropic.jl:92-106.
// in<0> = ε(6), in<1> = a3(3) (unit direction, the model's state); out<0> = σ(6), out<1> = E(36, optional)
//   E = L⊥ I⊗I + 2G⊥ Iˢʸᵐ + (L∥−L⊥)(I⊗A + A⊗I) + (M∥ − 4G∥ + 2G⊥ − 2L∥ + L⊥) A⊗A + 4(G∥−G⊥) 𝔸,   A = a⊗a,
//   𝔸_ijkl = ¼ (A_ik δ_jl + A_il δ_jk + δ_ik A_jl + δ_il A_jk)
// =================================================================================================
struct TransvIsoK { double L_perp, L_par, M_par, G_perp, G_par; };
MM_HD constexpr int slot_i(int c) { return c < 3 ? c : (c < 5 ? c - 2 : 2); }       // row index of storage slot c: 0,1,2,1,2,2
MM_HD constexpr int slot_j(int c) { return c < 3 ? 0 : (c < 5 ? 1 : 2); }           // column index:             0,0,0,1,1,2
MM_HD constexpr int sym_slot(int i, int j) { return i >= j ? (j == 0 ? i : (j == 1 ? i + 2 : 5)) : (i == 0 ? j : (i == 1 ? j + 2 : 5)); }
template <class Acc> MM_HD void transviso_point(const TransvIsoK& k, Acc& acc) {
    double e[6], a[3];
#pragma unroll
    for (int c = 0; c < 6; ++c) e[c] = acc.template in<0>(c);
#pragma unroll
    for (int c = 0; c < 3; ++c) a[c] = acc.template in<1>(c);
    double A[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) A[c] = a[slot_i(c)] * a[slot_j(c)];
    const double c3 = k.L_par - k.L_perp, c4 = k.M_par - 4.0 * k.G_par + 2.0 * k.G_perp - 2.0 * k.L_par + k.L_perp, c5 = k.G_par - k.G_perp;
    double c3A[6], c5A[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) { c3A[c] = c3 * A[c]; c5A[c] = c5 * A[c]; }
    double sig[6] = {0, 0, 0, 0, 0, 0};
    // the Kronecker deltas are compile-time constants after unrolling: only the surviving terms are emitted
#pragma unroll
    for (int J = 0; J < 6; ++J) {
        const double wJ = is_diag_slot(J) ? e[J] : 2.0 * e[J];      // ⊡ sums both (k,l) and (l,k)
#pragma unroll
        for (int I = 0; I < 6; ++I) {
            const int i = slot_i(I), j = slot_j(I), kk = slot_i(J), l = slot_j(J);
            double v = (c4 * A[I]) * A[J];
            if (i == j && kk == l) v += k.L_perp;
            if (i == kk && j == l) v += k.G_perp;
            if (i == l && j == kk) v += k.G_perp;
            if (i == j) v += c3A[J];
            if (kk == l) v += c3A[I];
            if (j == l) v += c5A[sym_slot(i, kk)];
            if (j == kk) v += c5A[sym_slot(i, l)];
            if (i == kk) v += c5A[sym_slot(j, l)];
            if (i == l) v += c5A[sym_slot(j, kk)];
            sig[I] = fma(v, wJ, sig[I]);
            if (acc.template has_out<1>()) acc.template out<1>(I + 6 * J, v);
        }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) acc.template out<0>(c, sig[c]);
}

// =================================================================================================
// elastic_strain_energy_density — LinearElastic.jl:63: ½ ε:Eᵉ:ε;
// neohook.jl:29-33, yeoh.jl:31-35, stvenant.jl:28-34 through any strain measure (traits.jl:133-136).
// in<0> = strain; out<0> = ψ (1 component)
// =================================================================================================
template <class Acc> MM_HD void linear_elastic_energy_point(const ElasticK& k, Acc& acc) {
    double e[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) e[c] = acc.template in<0>(c);
    const double tr = tr6(e);
    acc.template out<0>(0, 0.5 * (k.lam * tr * tr + 2.0 * k.G * ddot6(e, e)));
}
template <int MODEL, int KIND, class Acc> MM_HD void hyper_energy_point(const HyperK& k, Acc& acc) {
    double C[6];
    if (KIND == 1) {
        double F[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) F[c] = acc.template in<0>(c);
        C[0] = F[0] * F[0] + F[1] * F[1] + F[2] * F[2];
        C[1] = F[3] * F[0] + F[4] * F[1] + F[5] * F[2];
        C[2] = F[6] * F[0] + F[7] * F[1] + F[8] * F[2];
        C[3] = F[3] * F[3] + F[4] * F[4] + F[5] * F[5];
        C[4] = F[6] * F[3] + F[7] * F[4] + F[8] * F[5];
        C[5] = F[6] * F[6] + F[7] * F[7] + F[8] * F[8];
    } else if (KIND == 2) {
#pragma unroll
        for (int c = 0; c < 6; ++c) C[c] = is_diag_slot(c) ? (2.0 * acc.template in<0>(c) + 1.0) : (2.0 * acc.template in<0>(c));
    } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) C[c] = acc.template in<0>(c);
    }
    double psi;
    if (MODEL == 2) {      // StVenant: E = (C − I)/2, ψ = ½ λ tr(E)² + μ E:E
        double E[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) E[c] = 0.5 * (is_diag_slot(c) ? (C[c] - 1.0) : C[c]);
        const double trE = tr6(E);
        psi = 0.5 * k.lambda * (trE * trE) + k.mu * ddot6(E, E);
    } else {
        const double det = C[0] * (C[3] * C[5] - C[4] * C[4]) + C[1] * (C[2] * C[4] - C[1] * C[5]) + C[2] * (C[1] * C[4] - C[2] * C[3]);
        const double lJ = 0.5 * log(det), i3 = tr6(C) - 3.0;
        psi = k.mu / 2 * i3 - k.mu * lJ + k.lambda / 2 * (lJ * lJ);
        if (MODEL == 1) psi += k.c2 * (i3 * i3) + k.c3 * (i3 * i3 * i3);
    }
    acc.template out<0>(0, psi);
}

}  // namespace mm
