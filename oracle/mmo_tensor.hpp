// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product path;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Small fixed-size tensor algebra + forward-mode dual numbers, restating the pieces of the
// third-party Julia packages that the reference's hot path leans on (none of them is vendored
// under /root/reference; versions are only lower-bounded in reference Project.toml:14-21):
//   * Tensors.jl  >= 1.4   : SymmetricTensor{2,3}/{4,3}, Tensor{2,3}/{4,3}, dcontract, otimes, otimesu,
//                            otimesl, dev, norm, tr, det, inv, tdot, symmetric, tomandel/frommandel
//   * ForwardDiff >= 0.10  : Dual numbers (value + N partials), nested for Hessians
//   * LAPACK / StaticArrays: LU with partial pivoting for `\` and `inv`
// Storage conventions (pinned by the golden files):
//   SymmetricTensor{2,3}  -> 6 doubles, order (11,21,31,22,32,33)  [lower triangle, column-major]
//   Tensor{2,3}           -> 9 doubles, column-major (11,21,31,12,22,32,13,23,33)
//   SymmetricTensor{4,3}  -> 36 doubles, index I + 6*J, I<->(ij), J<->(kl), raw components
//   Mandel order          -> [11,22,33,23,13,12] with sqrt(2) on the shear entries
#pragma once
#include <cmath>
#include <cstddef>
#include <type_traits>

namespace mmo {

// ------------------------------------------------------------------------------------------
// Forward-mode dual numbers: value + N partials over a base field B (double, or another Dual
// for second derivatives — this is how Tensors.jl `hessian` nests ForwardDiff duals).
// ------------------------------------------------------------------------------------------
template <int N, class B = double>
struct Dual {
    B v;
    B d[N];
    Dual() : v(B(0.0)) { for (int i = 0; i < N; ++i) d[i] = B(0.0); }
    Dual(double c) : v(B(c)) { for (int i = 0; i < N; ++i) d[i] = B(0.0); }
    template <class B2 = B, class = typename std::enable_if<!std::is_same<B2, double>::value>::type>
    Dual(const B& c) : v(c) { for (int i = 0; i < N; ++i) d[i] = B(0.0); }
};

inline double value(double x) { return x; }
template <int N, class B> inline double value(const Dual<N, B>& x) { return value(x.v); }

// scalar helpers for plain doubles so templated code can call the same names
inline double sqrt_(double x) { return std::sqrt(x); }
inline double log_(double x) { return std::log(x); }
inline double exp_(double x) { return std::exp(x); }
inline double abs_(double x) { return std::fabs(x); }
inline double pow_(double x, double p) { return std::pow(x, p); }

#define MMO_T template <int N, class B>
#define MMO_D Dual<N, B>

MMO_T inline MMO_D operator+(const MMO_D& a, const MMO_D& b) { MMO_D r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
MMO_T inline MMO_D operator-(const MMO_D& a, const MMO_D& b) { MMO_D r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
MMO_T inline MMO_D operator-(const MMO_D& a) { MMO_D r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
// ForwardDiff product rule: partials = a.d*b.v + b.d*a.v
MMO_T inline MMO_D operator*(const MMO_D& a, const MMO_D& b) { MMO_D r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + b.d[i] * a.v; return r; }
// ForwardDiff quotient rule: (a.d*b.v - a.v*b.d) / b.v^2 written as a.d/b.v - (a.v/b.v^2) b.d
MMO_T inline MMO_D operator/(const MMO_D& a, const MMO_D& b) {
    MMO_D r; r.v = a.v / b.v; B inv = B(1.0) / b.v; B k = r.v * inv;
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * inv - k * b.d[i];
    return r;
}
MMO_T inline MMO_D operator+(const MMO_D& a, double c) { MMO_D r = a; r.v = a.v + c; return r; }
MMO_T inline MMO_D operator+(double c, const MMO_D& a) { return a + c; }
MMO_T inline MMO_D operator-(const MMO_D& a, double c) { MMO_D r = a; r.v = a.v - c; return r; }
MMO_T inline MMO_D operator-(double c, const MMO_D& a) { MMO_D r; r.v = c - a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
MMO_T inline MMO_D operator*(const MMO_D& a, double c) { MMO_D r; r.v = a.v * c; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * c; return r; }
MMO_T inline MMO_D operator*(double c, const MMO_D& a) { return a * c; }
MMO_T inline MMO_D operator/(const MMO_D& a, double c) { MMO_D r; r.v = a.v / c; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / c; return r; }
MMO_T inline MMO_D operator/(double c, const MMO_D& a) { return MMO_D(c) / a; }
MMO_T inline MMO_D& operator+=(MMO_D& a, const MMO_D& b) { a = a + b; return a; }
MMO_T inline MMO_D& operator-=(MMO_D& a, const MMO_D& b) { a = a - b; return a; }

MMO_T inline MMO_D sqrt_(const MMO_D& a) { MMO_D r; r.v = sqrt_(a.v); B k = B(0.5) / r.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * k; return r; }
MMO_T inline MMO_D log_(const MMO_D& a) { MMO_D r; r.v = log_(a.v); B k = B(1.0) / a.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * k; return r; }
MMO_T inline MMO_D exp_(const MMO_D& a) { MMO_D r; r.v = exp_(a.v); for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * r.v; return r; }
// ForwardDiff: abs(d) = signbit(value(d)) ? -d : d
MMO_T inline MMO_D abs_(const MMO_D& a) { return std::signbit(value(a)) ? -a : a; }
// ForwardDiff `^(::Dual, ::Real)`: value x^p, partials p*x^(p-1)*dx  (0 for x == 0, p > 1)
MMO_T inline MMO_D pow_(const MMO_D& a, double p) {
    MMO_D r; r.v = pow_(a.v, p); B k = p * pow_(a.v, p - 1.0);
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * k;
    return r;
}
#undef MMO_T
#undef MMO_D

// Julia `sign(x)`; ForwardDiff's `sign(::Dual)` returns the sign of the value (no partials).
template <class T> inline double sign_(const T& x) { double v = value(x); return (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : 0.0); }

// ------------------------------------------------------------------------------------------
// Tensors
// ------------------------------------------------------------------------------------------
// storage index of the symmetric pair (i,j), i,j in 0..2 : (11,21,31,22,32,33)
static const int SIDX[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
// the (i,j) of storage slot I
static const int SI[6] = {0, 1, 2, 1, 2, 2};
static const int SJ[6] = {0, 0, 0, 1, 1, 2};

template <class T> struct Sym2 {
    T a[6];
    Sym2() { for (int i = 0; i < 6; ++i) a[i] = T(0.0); }
    T& operator()(int i, int j) { return a[SIDX[i][j]]; }
    const T& operator()(int i, int j) const { return a[SIDX[i][j]]; }
};
template <class T> struct Ten2 {  // column-major 3x3
    T a[9];
    Ten2() { for (int i = 0; i < 9; ++i) a[i] = T(0.0); }
    T& operator()(int i, int j) { return a[i + 3 * j]; }
    const T& operator()(int i, int j) const { return a[i + 3 * j]; }
};
template <class T> struct Sym4 {  // minor-symmetric 4th order, 36 = I + 6 J
    T a[36];
    Sym4() { for (int i = 0; i < 36; ++i) a[i] = T(0.0); }
    T& operator()(int i, int j, int k, int l) { return a[SIDX[i][j] + 6 * SIDX[k][l]]; }
    const T& operator()(int i, int j, int k, int l) const { return a[SIDX[i][j] + 6 * SIDX[k][l]]; }
};
template <class T> struct Ten4 {  // full 4th order, 81 = i + 3j + 9k + 27l
    T a[81];
    Ten4() { for (int i = 0; i < 81; ++i) a[i] = T(0.0); }
    T& operator()(int i, int j, int k, int l) { return a[i + 3 * j + 9 * k + 27 * l]; }
    const T& operator()(int i, int j, int k, int l) const { return a[i + 3 * j + 9 * k + 27 * l]; }
};

template <class T> inline Sym2<T> identity2() { Sym2<T> I; I(0, 0) = T(1.0); I(1, 1) = T(1.0); I(2, 2) = T(1.0); return I; }

template <class T, class U> inline auto operator+(const Sym2<T>& x, const Sym2<U>& y) { Sym2<decltype(T() + U())> r; for (int i = 0; i < 6; ++i) r.a[i] = x.a[i] + y.a[i]; return r; }
template <class T, class U> inline auto operator-(const Sym2<T>& x, const Sym2<U>& y) { Sym2<decltype(T() - U())> r; for (int i = 0; i < 6; ++i) r.a[i] = x.a[i] - y.a[i]; return r; }
template <class T, class S> inline auto scale(const S& s, const Sym2<T>& x) { Sym2<decltype(S() * T())> r; for (int i = 0; i < 6; ++i) r.a[i] = s * x.a[i]; return r; }

template <class T> inline T tr(const Sym2<T>& x) { return x(0, 0) + x(1, 1) + x(2, 2); }
// Tensors.jl dev(A) = A - tr(A)/3 * I
template <class T> inline Sym2<T> dev(const Sym2<T>& x) { Sym2<T> r = x; T m = tr(x) / 3.0; r(0, 0) = x(0, 0) - m; r(1, 1) = x(1, 1) - m; r(2, 2) = x(2, 2) - m; return r; }
// full double contraction A:B over all 9 (i,j)
template <class T, class U> inline auto dcontract(const Sym2<T>& x, const Sym2<U>& y) {
    decltype(T() * U()) s(0.0);
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) s = s + x(i, j) * y(i, j);
    return s;
}
template <class T, class U> inline auto dcontract(const Sym2<T>& x, const Ten2<U>& y) {
    decltype(T() * U()) s(0.0);
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) s = s + x(i, j) * y(i, j);
    return s;
}
// Frobenius norm over all 9 entries (Tensors.jl norm)
template <class T> inline T norm(const Sym2<T>& x) { return sqrt_(dcontract(x, x)); }

// E : e   (SymmetricTensor{4} ⊡ SymmetricTensor{2}), sum over all 9 (k,l)
template <class T, class U> inline auto dcontract(const Sym4<T>& E, const Sym2<U>& e) {
    Sym2<decltype(T() * U())> r;
    for (int I = 0; I < 6; ++I) {
        decltype(T() * U()) s(0.0);
        for (int l = 0; l < 3; ++l) for (int k = 0; k < 3; ++k) s = s + E.a[I + 6 * SIDX[k][l]] * e(k, l);
        r.a[I] = s;
    }
    return r;
}
// E : M with a full (possibly unsymmetric) second-order tensor; E minor-symmetric -> symmetric result
template <class T, class U> inline auto dcontract(const Sym4<T>& E, const Ten2<U>& m) {
    Sym2<decltype(T() * U())> r;
    for (int I = 0; I < 6; ++I) {
        decltype(T() * U()) s(0.0);
        for (int l = 0; l < 3; ++l) for (int k = 0; k < 3; ++k) s = s + E.a[I + 6 * SIDX[k][l]] * m(k, l);
        r.a[I] = s;
    }
    return r;
}
// A ⊡ B for two minor-symmetric 4th-order tensors: (A:B)_ijmn = sum_kl A_ijkl B_klmn (all 9 kl)
template <class T> inline Sym4<T> dcontract(const Sym4<T>& A, const Sym4<T>& Bt) {
    Sym4<T> r;
    for (int J = 0; J < 6; ++J) for (int I = 0; I < 6; ++I) {
        T s(0.0);
        for (int l = 0; l < 3; ++l) for (int k = 0; k < 3; ++k) s = s + A.a[I + 6 * SIDX[k][l]] * Bt.a[SIDX[k][l] + 6 * J];
        r.a[I + 6 * J] = s;
    }
    return r;
}

template <class T> inline T det(const Sym2<T>& c) {
    return c(0, 0) * (c(1, 1) * c(2, 2) - c(1, 2) * c(2, 1)) - c(0, 1) * (c(1, 0) * c(2, 2) - c(1, 2) * c(2, 0)) +
           c(0, 2) * (c(1, 0) * c(2, 1) - c(1, 1) * c(2, 0));
}
// Tensors.jl inv for 3x3: adjugate / det
template <class T> inline Sym2<T> inv(const Sym2<T>& c) {
    T idet = T(1.0) / det(c);
    Sym2<T> r;
    r(0, 0) = (c(1, 1) * c(2, 2) - c(1, 2) * c(2, 1)) * idet;
    r(1, 0) = (c(1, 2) * c(2, 0) - c(1, 0) * c(2, 2)) * idet;
    r(2, 0) = (c(1, 0) * c(2, 1) - c(1, 1) * c(2, 0)) * idet;
    r(1, 1) = (c(0, 0) * c(2, 2) - c(0, 2) * c(2, 0)) * idet;
    r(2, 1) = (c(0, 1) * c(2, 0) - c(0, 0) * c(2, 1)) * idet;
    r(2, 2) = (c(0, 0) * c(1, 1) - c(0, 1) * c(1, 0)) * idet;
    return r;
}
// tdot(F) = Fᵀ·F  (reference src/traits.jl:36)
template <class T> inline Sym2<T> tdot(const Ten2<T>& F) {
    Sym2<T> C;
    for (int j = 0; j < 3; ++j) for (int i = j; i < 3; ++i) {
        T s(0.0);
        for (int k = 0; k < 3; ++k) s = s + F(k, i) * F(k, j);
        C(i, j) = s;
    }
    return C;
}

// 4th-order products on full tensors
template <class T> inline Ten4<T> otimes(const Sym2<T>& A, const Sym2<T>& Bt) { Ten4<T> r; for (int l = 0; l < 3; ++l) for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) r(i, j, k, l) = A(i, j) * Bt(k, l); return r; }
// otimesu(A,B)_ijkl = A_ik B_jl ; otimesl(A,B)_ijkl = A_il B_jk
template <class T> inline Ten4<T> otimesu(const Sym2<T>& A, const Sym2<T>& Bt) { Ten4<T> r; for (int l = 0; l < 3; ++l) for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) r(i, j, k, l) = A(i, k) * Bt(j, l); return r; }
template <class T> inline Ten4<T> otimesl(const Sym2<T>& A, const Sym2<T>& Bt) { Ten4<T> r; for (int l = 0; l < 3; ++l) for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) r(i, j, k, l) = A(i, l) * Bt(j, k); return r; }
template <class T> inline Ten4<T> operator+(const Ten4<T>& x, const Ten4<T>& y) { Ten4<T> r; for (int i = 0; i < 81; ++i) r.a[i] = x.a[i] + y.a[i]; return r; }
template <class T, class S> inline Ten4<T> scale(const S& s, const Ten4<T>& x) { Ten4<T> r; for (int i = 0; i < 81; ++i) r.a[i] = s * x.a[i]; return r; }
// Tensors.jl symmetric(::Tensor{4}) = minor symmetrisation
template <class T> inline Sym4<T> symmetric(const Ten4<T>& A) {
    Sym4<T> r;
    for (int J = 0; J < 6; ++J) for (int I = 0; I < 6; ++I) {
        int i = SI[I], j = SJ[I], k = SI[J], l = SJ[J];
        r.a[I + 6 * J] = (A(i, j, k, l) + A(j, i, k, l) + A(i, j, l, k) + A(j, i, l, k)) / 4.0;
    }
    return r;
}

// ------------------------------------------------------------------------------------------
// Dense LU with partial pivoting (LAPACK dgetrf / StaticArrays lu semantics), column-major n x n.
// Returns false on an exactly singular pivot.
// ------------------------------------------------------------------------------------------
template <int NN> inline bool lu_factor(double* A, int* piv) {
    for (int k = 0; k < NN; ++k) {
        int p = k; double amax = std::fabs(A[k + NN * k]);
        for (int i = k + 1; i < NN; ++i) { double v = std::fabs(A[i + NN * k]); if (v > amax) { amax = v; p = i; } }
        piv[k] = p;
        if (amax == 0.0) return false;
        if (p != k) for (int j = 0; j < NN; ++j) { double t = A[k + NN * j]; A[k + NN * j] = A[p + NN * j]; A[p + NN * j] = t; }
        double ipiv = 1.0 / A[k + NN * k];
        for (int i = k + 1; i < NN; ++i) A[i + NN * k] *= ipiv;
        for (int j = k + 1; j < NN; ++j) { double akj = A[k + NN * j]; for (int i = k + 1; i < NN; ++i) A[i + NN * j] -= A[i + NN * k] * akj; }
    }
    return true;
}
template <int NN> inline void lu_solve(const double* LU, const int* piv, double* b) {
    for (int k = 0; k < NN; ++k) { int p = piv[k]; if (p != k) { double t = b[k]; b[k] = b[p]; b[p] = t; } }
    for (int k = 0; k < NN; ++k) for (int i = k + 1; i < NN; ++i) b[i] -= LU[i + NN * k] * b[k];
    for (int k = NN - 1; k >= 0; --k) { b[k] /= LU[k + NN * k]; for (int i = 0; i < k; ++i) b[i] -= LU[i + NN * k] * b[k]; }
}
template <int NN> inline bool lu_inverse(const double* A, double* Ainv) {
    double LU[NN * NN]; int piv[NN];
    for (int i = 0; i < NN * NN; ++i) LU[i] = A[i];
    if (!lu_factor<NN>(LU, piv)) return false;
    for (int j = 0; j < NN; ++j) {
        double e[NN]; for (int i = 0; i < NN; ++i) e[i] = (i == j) ? 1.0 : 0.0;
        lu_solve<NN>(LU, piv, e);
        for (int i = 0; i < NN; ++i) Ainv[i + NN * j] = e[i];
    }
    return true;
}

// Mandel maps for SymmetricTensor{2,3}: [11,22,33,23,13,12], sqrt(2) on shear
static const int MANDEL_I[6] = {0, 1, 2, 1, 0, 0};
static const int MANDEL_J[6] = {0, 1, 2, 2, 2, 1};
template <class T> inline void tomandel(T* v, const Sym2<T>& s) { const double r2 = std::sqrt(2.0); for (int m = 0; m < 6; ++m) v[m] = (m < 3) ? s(MANDEL_I[m], MANDEL_J[m]) : s(MANDEL_I[m], MANDEL_J[m]) * r2; }
template <class T> inline Sym2<T> frommandel(const T* v) { const double r2 = std::sqrt(2.0); Sym2<T> s; for (int m = 0; m < 6; ++m) s(MANDEL_I[m], MANDEL_J[m]) = (m < 3) ? v[m] : v[m] / r2; return s; }
// Mandel order for a full Tensor{2,3}: [11,22,33,23,13,12,32,31,21]
static const int MANDEL9_I[9] = {0, 1, 2, 1, 0, 0, 2, 2, 1};
static const int MANDEL9_J[9] = {0, 1, 2, 2, 2, 1, 1, 0, 0};

}  // namespace mmo
