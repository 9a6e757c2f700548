// Dimension wrappers — reference src/wrappers.jl:1-162, fused around the 3D point functions of mm_math.cuh.
//   strain-restricted (UniaxialStrain{1}, PlaneStrain{2}): pad the strain with zeros, one 3D evaluation, cut σ and ∂σ∂ε
//   down to dimension d (:29-41);
//   stress-restricted (UniaxialStress{1}, PlaneStress{2}, ShellStress{3}): outer Newton on the out-of-plane strain
//   components until the out-of-plane stresses vanish (2-norm of their Mandel vector < tol, default 1e-8, max_iter 10),
//   tangent = Schur complement of the Mandel 6x6 (:45-79).
// The inner model runs on a REGISTER accessor (RegIO): its σ, ∂σ∂ε and new state never leave the thread between outer
// iterations; only the reduced quantities are stored.  Reduced storage = Tensors.jl SymmetricTensor{2,d}/{4,d}:
// d = 1: (11);  d = 2: (11,21,22), 9 = I + 3J;  ShellStress keeps the full 3D storage with zero 33-rows/columns (:131-146).
#pragma once
#include "mm_math.cuh"

namespace mm {

enum { WRAP_UNIAXIAL_STRAIN = 1, WRAP_PLANE_STRAIN = 2, WRAP_UNIAXIAL_STRESS = 3, WRAP_PLANE_STRESS = 4, WRAP_SHELL_STRESS = 5 };

MM_HD constexpr int wrap_dim(int kind) { return (kind == WRAP_UNIAXIAL_STRAIN || kind == WRAP_UNIAXIAL_STRESS) ? 1 : ((kind == WRAP_SHELL_STRESS) ? 3 : 2); }
MM_HD constexpr int wrap_nc(int kind) { return wrap_dim(kind) * (wrap_dim(kind) + 1) / 2; }
// 3D storage slot (11,21,31,22,32,33) of reduced storage slot r
MM_HD constexpr int wrap_slot(int d, int r) { return d == 1 ? 0 : (d == 2 ? (r == 0 ? 0 : (r == 1 ? 1 : 3)) : r); }
// 3D storage slot of Mandel position [11,22,33,23,13,12]
MM_HD constexpr int slot_of_mandel(int m) { return m == 0 ? 0 : (m == 1 ? 3 : (m == 2 ? 5 : (m == 3 ? 4 : (m == 4 ? 2 : 1)))); }
// Mandel index sets of the restricted-stress kinds (wrappers.jl:149-162, 0-based)
MM_HD constexpr int wrap_nzero(int kind) { return kind == WRAP_PLANE_STRESS ? 3 : (kind == WRAP_UNIAXIAL_STRESS ? 5 : 1); }
MM_HD constexpr int wrap_zero(int kind, int a) { return kind == WRAP_PLANE_STRESS ? 2 + a : (kind == WRAP_UNIAXIAL_STRESS ? 1 + a : 2); }
MM_HD constexpr int wrap_nz(int kind, int a) {
    return kind == WRAP_PLANE_STRESS ? (a == 0 ? 0 : (a == 1 ? 1 : 5)) : (kind == WRAP_UNIAXIAL_STRESS ? 0 : (a < 2 ? a : a + 1));
}
// reduced storage slot of the a-th non-zero Mandel position
MM_HD constexpr int wrap_red_of_nz(int kind, int a) {
    return kind == WRAP_PLANE_STRESS ? (a == 0 ? 0 : (a == 1 ? 2 : 1)) : (kind == WRAP_UNIAXIAL_STRESS ? 0 : slot_of_mandel(wrap_nz(kind, a)));
}
MM_HD constexpr double mandel_f(int m) { return m < 3 ? 1.0 : 1.4142135623730950488; }

// register-resident accessor for one inner 3D evaluation (NS = number of state doubles of the inner model)
template <int NS> struct RegIO {
    double eps[6];
    double st_in[NS > 0 ? NS : 1];
    double sig[6], tan[36];
    double st_out[NS > 0 ? NS : 1];
    template <int A> MM_HD double in(int c) const { return A == 0 ? eps[c] : st_in[c]; }
    template <int A> MM_HD void out(int c, double v) { if (A == 0) sig[c] = v; else if (A == 1) tan[c] = v; else st_out[c] = v; }
    template <int A> MM_HD bool has_out() const { return true; }
    MM_HD double T(int idx) const { return tan[idx]; }          // tangent entry I + 6 J as the wrapper reads it (a lazy IO computes it on demand)
};

// solve A X = B in place (A: n x n row-major, B: n x m row-major) by Gaussian elimination with partial pivoting;
// everything statically indexed so it stays in registers.  false on a zero pivot.
// The elimination column k is a TEMPLATE parameter: with a plain `#pragma unroll` over k nvcc left the pivot search of the 5 x 5 case
// (UniaxialStress) as a run-time loop, which put A and B into local memory (346 LDL/STL in the kernel, 248 B of stack).
template <int n, int m, int k> struct SmallElim {
    static MM_HD void run(double* A, double* B, double* ip, bool& ok) {
#pragma unroll
        for (int i = k + 1; i < n; ++i) {            // bring the largest |A[i][k]| to row k
            if (fabs(A[i * n + k]) > fabs(A[k * n + k])) {
#pragma unroll
                for (int j = k; j < n; ++j) { const double t = A[k * n + j]; A[k * n + j] = A[i * n + j]; A[i * n + j] = t; }
#pragma unroll
                for (int j = 0; j < m; ++j) { const double t = B[k * m + j]; B[k * m + j] = B[i * m + j]; B[i * m + j] = t; }
            }
        }
        const double pv = A[k * n + k];
        ok = ok & (pv != 0.0);
        ip[k] = fast_rcp(pv);                        // pivot reciprocals, reused by the back substitution (no IEEE division scaffolding)
#pragma unroll
        for (int i = k + 1; i < n; ++i) {
            const double l = A[i * n + k] * ip[k];
#pragma unroll
            for (int j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
#pragma unroll
            for (int j = 0; j < m; ++j) B[i * m + j] -= l * B[k * m + j];
        }
        SmallElim<n, m, k + 1>::run(A, B, ip, ok);
    }
};
template <int n, int m> struct SmallElim<n, m, n> { static MM_HD void run(double*, double*, double*, bool&) {} };
template <int n, int m, int k> struct SmallBack {
    static MM_HD void run(const double* A, double* B, const double* ip) {
#pragma unroll
        for (int j = 0; j < m; ++j) {
            double s = B[k * m + j];
#pragma unroll
            for (int i = k + 1; i < n; ++i) s -= A[k * n + i] * B[i * m + j];
            B[k * m + j] = s * ip[k];
        }
        SmallBack<n, m, k - 1>::run(A, B, ip);
    }
};
template <int n, int m> struct SmallBack<n, m, -1> { static MM_HD void run(const double*, double*, const double*) {} };
template <int n, int m> MM_HD bool small_solve(double* A, double* B) {
    bool ok = true;
    double ip[n];
    SmallElim<n, m, 0>::run(A, B, ip, ok);
    SmallBack<n, m, n - 1>::run(A, B, ip);
    return ok;
}

// status bits returned: 0x80 inner model failed, 0x40 outer Newton not converged ("Could not find plane stress state")
// inner(io, &inner_iters) -> inner flag bits (0x80 = failed).  acc: in<0> reduced strain, in<1> state (NS > 0),
// out<0> reduced stress, out<1> reduced tangent (optional), out<2> new state (NS > 0).
// wrapped_core: everything but the state traffic.  `io` is what the inner 3D model sees: eps[6] in, sig[6] out as members, T(I + 6 J)
// the tangent entries (RegIO: the stored 36; the J2 wrapper: computed on demand from the model's tangent ingredients).
template <int KIND, class IO, class Inner, class Acc>
MM_HD int wrapped_core(Inner&& inner, Acc& acc, IO& io, double tol, int max_iter, int* outer_iters, int* inner_iters) {
    constexpr int d = wrap_dim(KIND), nc = wrap_nc(KIND);
#pragma unroll
    for (int c = 0; c < 6; ++c) io.eps[c] = 0.0;
#pragma unroll
    for (int r = 0; r < nc; ++r) io.eps[wrap_slot(d, r)] = acc.template in<0>(r);               // increase_dim (:24-27)
    int flag = 0, oit = 0;
    if (KIND == WRAP_UNIAXIAL_STRAIN || KIND == WRAP_PLANE_STRAIN) {                              // :29-41
        flag = inner(io, inner_iters);
#pragma unroll
        for (int r = 0; r < nc; ++r) acc.template out<0>(r, io.sig[wrap_slot(d, r)]);
        if (acc.template has_out<1>()) {
#pragma unroll
            for (int b = 0; b < nc; ++b)
#pragma unroll
                for (int a = 0; a < nc; ++a) acc.template out<1>(a + nc * b, io.T(wrap_slot(d, a) + 6 * wrap_slot(d, b)));
        }
    } else {
        constexpr int nzero = wrap_nzero(KIND), nnz = 6 - nzero;
        bool done = false;
        for (int it = 1; it <= max_iter && !done; ++it) {                                         // :60-77
            flag = inner(io, inner_iters);
            if (flag & 0x80) break;
            oit = it - 1;
            double sm[nzero], nrm2 = 0.0, Jzz[nzero * nzero];
#pragma unroll
            for (int a = 0; a < nzero; ++a) {
                sm[a] = io.sig[slot_of_mandel(wrap_zero(KIND, a))] * mandel_f(wrap_zero(KIND, a));
                nrm2 += sm[a] * sm[a];
#pragma unroll
                for (int b = 0; b < nzero; ++b)
                    Jzz[a * nzero + b] = io.T(slot_of_mandel(wrap_zero(KIND, a)) + 6 * slot_of_mandel(wrap_zero(KIND, b))) *
                                         (mandel_f(wrap_zero(KIND, a)) * mandel_f(wrap_zero(KIND, b)));
            }
            if (sqrt(nrm2) < tol) {                                                               // :63-70 Schur complement
                double X[nzero * nnz];                                                            // inv(M_zz) M_z,nz
#pragma unroll
                for (int a = 0; a < nzero; ++a)
#pragma unroll
                    for (int b = 0; b < nnz; ++b)
                        X[a * nnz + b] = io.T(slot_of_mandel(wrap_zero(KIND, a)) + 6 * slot_of_mandel(wrap_nz(KIND, b))) *
                                         (mandel_f(wrap_zero(KIND, a)) * mandel_f(wrap_nz(KIND, b)));
                if (!small_solve<nzero, nnz>(Jzz, X)) flag |= 0x40;
#pragma unroll
                for (int r = 0; r < nc; ++r) acc.template out<0>(r, io.sig[wrap_slot(d, r)]);
                if (acc.template has_out<1>()) {
                    if (KIND == WRAP_SHELL_STRESS) {
#pragma unroll
                        for (int c = 0; c < 36; ++c) if (c % 6 == 5 || c / 6 == 5) acc.template out<1>(c, 0.0);   // zero 33-row/column
                    }
#pragma unroll
                    for (int a = 0; a < nnz; ++a)
#pragma unroll
                        for (int b = 0; b < nnz; ++b) {
                            const double fa = mandel_f(wrap_nz(KIND, a)), fb = mandel_f(wrap_nz(KIND, b));
                            double v = io.T(slot_of_mandel(wrap_nz(KIND, a)) + 6 * slot_of_mandel(wrap_nz(KIND, b))) * (fa * fb);
#pragma unroll
                            for (int p = 0; p < nzero; ++p)
                                v -= io.T(slot_of_mandel(wrap_nz(KIND, a)) + 6 * slot_of_mandel(wrap_zero(KIND, p))) * (fa * mandel_f(wrap_zero(KIND, p))) * X[p * nnz + b];
                            acc.template out<1>(wrap_red_of_nz(KIND, a) + nc * wrap_red_of_nz(KIND, b), v * (1.0 / (fa * fb)));   // constant folded
                        }
                }
                done = true;
            } else {                                                                              // :72-76 Δε_3D += frommandel(−inv(J) σ)
                if (!small_solve<nzero, 1>(Jzz, sm)) { flag |= 0x40; break; }
#pragma unroll
                for (int a = 0; a < nzero; ++a) io.eps[slot_of_mandel(wrap_zero(KIND, a))] -= sm[a] * (1.0 / mandel_f(wrap_zero(KIND, a)));   // constant folded
                oit = it;
            }
        }
        if (!done && !(flag & 0x80)) flag |= 0x40;                                                // :78
        if (!done) {                                                                              // failed: still define the outputs
#pragma unroll
            for (int r = 0; r < nc; ++r) acc.template out<0>(r, io.sig[wrap_slot(d, r)]);
            if (acc.template has_out<1>()) {
#pragma unroll
                for (int c = 0; c < nc * nc; ++c) acc.template out<1>(c, 0.0);
            }
        }
    }
    *outer_iters = oit;
    return flag;
}
// ---- the two halves of one outer iteration as stand-alone steps (same expressions as wrapped_core above), for the kernels that do
// not run the outer loop inside one thread: the J2 state machine (mm_wrap_lp.cuh) and the multi-launch crystal wrapper (mm_wrap.cu).
// T(I, J) = tangent entry ∂σ_I/∂ε_J of the last 3D evaluation in raw storage slots.
// wrap_outer_step: 0 = out-of-plane stresses below tol (converged), 1 = eps updated by one Newton step, 2 = singular zero-stress block
template <int KIND, class TF> MM_HD int wrap_outer_step(const TF& T, const double* sig, double* eps, double tol) {
    constexpr int nzero = wrap_nzero(KIND);
    double sm[nzero], nrm2 = 0.0, Jzz[nzero * nzero];
#pragma unroll
    for (int a = 0; a < nzero; ++a) {
        sm[a] = sig[slot_of_mandel(wrap_zero(KIND, a))] * mandel_f(wrap_zero(KIND, a));
        nrm2 += sm[a] * sm[a];
    }
    if (sqrt(nrm2) < tol) return 0;                                                           // wrappers.jl:63
#pragma unroll
    for (int a = 0; a < nzero; ++a)
#pragma unroll
        for (int b = 0; b < nzero; ++b)
            Jzz[a * nzero + b] = T(slot_of_mandel(wrap_zero(KIND, a)), slot_of_mandel(wrap_zero(KIND, b))) * (mandel_f(wrap_zero(KIND, a)) * mandel_f(wrap_zero(KIND, b)));
    if (!small_solve<nzero, 1>(Jzz, sm)) return 2;
#pragma unroll
    for (int a = 0; a < nzero; ++a) eps[slot_of_mandel(wrap_zero(KIND, a))] -= sm[a] * (1.0 / mandel_f(wrap_zero(KIND, a)));   // :72-76
    return 1;
}
// wrap_schur_out: reduced tangent = Schur complement of the Mandel 6x6 (wrappers.jl:63-70) into out<1> (if requested); false on a
// singular zero-stress block (the solve runs either way: it decides the 0x40 flag)
template <int KIND, class TF, class Acc> MM_HD bool wrap_schur_out(const TF& T, Acc& acc) {
    constexpr int nc = wrap_nc(KIND), nzero = wrap_nzero(KIND), nnz = 6 - nzero;
    double Jzz[nzero * nzero], X[nzero * nnz];                     // X = inv(M_zz) M_z,nz
#pragma unroll
    for (int a = 0; a < nzero; ++a) {
#pragma unroll
        for (int b = 0; b < nzero; ++b)
            Jzz[a * nzero + b] = T(slot_of_mandel(wrap_zero(KIND, a)), slot_of_mandel(wrap_zero(KIND, b))) * (mandel_f(wrap_zero(KIND, a)) * mandel_f(wrap_zero(KIND, b)));
#pragma unroll
        for (int b = 0; b < nnz; ++b)
            X[a * nnz + b] = T(slot_of_mandel(wrap_zero(KIND, a)), slot_of_mandel(wrap_nz(KIND, b))) * (mandel_f(wrap_zero(KIND, a)) * mandel_f(wrap_nz(KIND, b)));
    }
    const bool ok = small_solve<nzero, nnz>(Jzz, X);
    if (acc.template has_out<1>()) {
        if (KIND == WRAP_SHELL_STRESS) {
#pragma unroll
            for (int c = 0; c < 36; ++c) if (c % 6 == 5 || c / 6 == 5) acc.template out<1>(c, 0.0);   // zero 33-row/column
        }
#pragma unroll
        for (int a = 0; a < nnz; ++a)
#pragma unroll
            for (int b = 0; b < nnz; ++b) {
                const double fa = mandel_f(wrap_nz(KIND, a)), fb = mandel_f(wrap_nz(KIND, b));
                double v = T(slot_of_mandel(wrap_nz(KIND, a)), slot_of_mandel(wrap_nz(KIND, b))) * (fa * fb);
#pragma unroll
                for (int p = 0; p < nzero; ++p)
                    v -= T(slot_of_mandel(wrap_nz(KIND, a)), slot_of_mandel(wrap_zero(KIND, p))) * (fa * mandel_f(wrap_zero(KIND, p))) * X[p * nnz + b];
                acc.template out<1>(wrap_red_of_nz(KIND, a) + nc * wrap_red_of_nz(KIND, b), v * (1.0 / (fa * fb)));   // constant folded
            }
    }
    return ok;
}

template <int KIND, int NS, class Inner, class Acc>
MM_HD int wrapped_point(Inner&& inner, Acc& acc, double tol, int max_iter, int* outer_iters, int* inner_iters) {
    RegIO<NS> io;
#pragma unroll
    for (int c = 0; c < NS; ++c) io.st_in[c] = acc.template in<1>(c);
    const int flag = wrapped_core<KIND>(inner, acc, io, tol, max_iter, outer_iters, inner_iters);
#pragma unroll
    for (int c = 0; c < NS; ++c) acc.template out<2>(c, io.st_out[c]);
    return flag;
}

}  // namespace mm
