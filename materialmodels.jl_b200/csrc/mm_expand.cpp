// Host side of the compact J2 return (mm_compact.cu / mmh_plastic_compact): rebuild the dense per-point arrays an existing caller
// expects — ∂σ∂ε (36) and the new state (14) for all N points — from the per-yielded-point rows.  Reference semantics being undone:
// src/Plastic.jl:134-135 hands every elastic point the one stored m.Eᵉ and its unchanged state; this writes those copies out.
// Plain C++ (no CUDA): a memory-bound scatter, split over host threads by point range; the row index of a point is the number of
// yielded points before it, so each thread first counts its range (exclusive scan over the ranges), then expands it.
// The tangent of a factored row is evaluated with plastic_tangent_entry_lg — the very fma sequence of the device kernels — so the
// expanded arrays are bit-identical to what the dense entry point returns.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define MM_NT_STORES 1
#endif
#include "../../include/mmb200.h"
#include "mm_math.cuh"

namespace mm {
int set_error(int code, const char* fmt, ...);

namespace {
int g_host_threads = 0;

int host_threads() {
    if (g_host_threads > 0) return g_host_threads;
    unsigned h = std::thread::hardware_concurrency();
    if (h == 0) h = 4;
    return (int)(h > 64 ? 64 : h);
}

struct ExpandJob {
    int row_kind, layout;
    const uint8_t* flags;
    const double* rows;
    const double* state_in;
    double* tan;
    double* state_out;
    int64_t N;
    int64_t flags_base;   // global index of flags[0] (0 for a whole-array job; i0 for a chunk whose flags sit in a staging buffer)
    double lam, G;
    double Ee[36];
};

constexpr int BLK = 256;

// The expanded arrays are written once and not read here: non-temporal 8-byte stores (MOVNTI; any 8-byte alignment, full lines are
// combined in the write-combining buffers because every stream below is written in runs of >= 2 KB) skip the read-for-ownership of a
// normal store — the scatter is bound by host memory traffic, which this halves (tools/expand_bench.py).
static inline void put(double* dst, double v) {
#ifdef MM_NT_STORES
    long long b;
    memcpy(&b, &v, 8);
    _mm_stream_si64(reinterpret_cast<long long*>(dst), b);
#else
    *dst = v;
#endif
}
static inline void put_n(double* dst, const double* src, int n) { for (int i = 0; i < n; ++i) put(dst + i, src[i]); }
static inline void store_fence() {
#ifdef MM_NT_STORES
    _mm_sfence();
#endif
}

// one SoA component run of the factored tangent for a block of points: dst[t] = fma(u[t], n[t], fma(-g[t], idev, iso)) — the device
// kernels' expression (plastic_tangent_entry_lg), so the bits are theirs; elastic points carry g = u = n = 0 and get Eᵉ exactly.
static void tangent_run_scalar(double* dst, const double* u, const double* n, const double* g, double idev, double iso, int nb) {
    for (int t = 0; t < nb; ++t) put(dst + t, fma(u[t], n[t], fma(-g[t], idev, iso)));
}
#ifdef MM_NT_STORES
__attribute__((target("avx2,fma")))
static void tangent_run_avx2(double* dst, const double* u, const double* n, const double* g, double idev, double iso, int nb) {
    int t = 0;
    for (; t < nb && (reinterpret_cast<uintptr_t>(dst + t) & 31u); ++t) put(dst + t, fma(u[t], n[t], fma(-g[t], idev, iso)));
    const __m256d vid = _mm256_set1_pd(idev), vis = _mm256_set1_pd(iso);
    for (; t + 4 <= nb; t += 4) {
        const __m256d b = _mm256_fnmadd_pd(_mm256_loadu_pd(g + t), vid, vis);                       // fma(-g, idev, iso)
        _mm256_stream_pd(dst + t, _mm256_fmadd_pd(_mm256_loadu_pd(u + t), _mm256_loadu_pd(n + t), b));
    }
    for (; t < nb; ++t) put(dst + t, fma(u[t], n[t], fma(-g[t], idev, iso)));
}
#endif
typedef void (*tangent_run_fn)(double*, const double*, const double*, const double*, double, double, int);
static tangent_run_fn pick_tangent_run() {
#ifdef MM_NT_STORES
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return tangent_run_avx2;
#endif
    return tangent_run_scalar;
}

// points [lo, hi), first row index `row0`
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("fma")))
#endif
void expand_range(const ExpandJob& j, int64_t lo, int64_t hi, int64_t row0) {
    const bool factored = (j.row_kind == MM_ROWS_FACTORED_STATE || j.row_kind == MM_ROWS_FACTORED);
    const bool with_state = (j.row_kind == MM_ROWS_TANGENT_STATE || j.row_kind == MM_ROWS_FACTORED_STATE);
    const int TW = factored ? 13 : 36, W = TW + (with_state ? 14 : 0);
    const bool soa = j.layout == MM_LAYOUT_SOA;
    const bool inplace = (j.state_out == j.state_in);
    const int64_t N = j.N;
    int64_t ridx[BLK];
    int64_t row = row0;
    for (int64_t b0 = lo; b0 < hi; b0 += BLK) {
        const int nb = (int)((hi - b0 < BLK) ? (hi - b0) : BLK);
        for (int t = 0; t < nb; ++t) ridx[t] = (j.flags[b0 + t - j.flags_base] & MM_FLAG_PLASTIC) ? row++ : -1;
        if (j.tan && soa && factored) {
            // gather the block's factored rows into component-major runs (zeros for the elastic points), then 36 straight vector loops
            alignas(64) double gb[BLK], ub[6][BLK], nbuf[6][BLK];
            for (int t = 0; t < nb; ++t) {
                if (ridx[t] < 0) { gb[t] = 0.0; for (int c = 0; c < 6; ++c) { ub[c][t] = 0.0; nbuf[c][t] = 0.0; } continue; }
                const double* r = j.rows + ridx[t] * W;
                gb[t] = r[0];
                for (int c = 0; c < 6; ++c) { ub[c][t] = r[1 + c]; nbuf[c][t] = r[7 + c]; }
            }
            static const tangent_run_fn run = pick_tangent_run();
            for (int J = 0; J < 6; ++J)
                for (int I = 0; I < 6; ++I)
                    run(j.tan + (int64_t)(I + 6 * J) * N + b0, ub[I], nbuf[J], gb, idev_entry(I, J), iso_entry(I, J, j.lam, j.G), nb);
        } else if (j.tan) {
            if (soa) {
                for (int J = 0; J < 6; ++J)
                    for (int I = 0; I < 6; ++I) {
                        const int c = I + 6 * J;
                        double* dst = j.tan + (int64_t)c * N + b0;
                        const double e = j.Ee[c];
                        if (factored) {
                            for (int t = 0; t < nb; ++t) {
                                if (ridx[t] < 0) { put(dst + t, e); continue; }
                                const double* r = j.rows + ridx[t] * W;
                                put(dst + t, plastic_tangent_entry_lg(j.lam, j.G, r[0], r + 1, r + 7, I, J));
                            }
                        } else {
                            for (int t = 0; t < nb; ++t) put(dst + t, ridx[t] < 0 ? e : j.rows[ridx[t] * W + c]);
                        }
                    }
            } else {
                for (int t = 0; t < nb; ++t) {
                    double* dst = j.tan + (b0 + t) * 36;
                    if (ridx[t] < 0) { put_n(dst, j.Ee, 36); continue; }
                    const double* r = j.rows + ridx[t] * W;
                    if (factored) {
                        for (int J = 0; J < 6; ++J)
                            for (int I = 0; I < 6; ++I) put(dst + I + 6 * J, plastic_tangent_entry_lg(j.lam, j.G, r[0], r + 1, r + 7, I, J));
                    } else {
                        put_n(dst, r, 36);
                    }
                }
            }
        }
        if (j.state_out && with_state) {
            if (soa) {
                for (int c = 0; c < 14; ++c) {
                    double* dst = j.state_out + (int64_t)c * N + b0;
                    const double* src = j.state_in + (int64_t)c * N + b0;
                    for (int t = 0; t < nb; ++t) {
                        if (inplace) { if (ridx[t] >= 0) dst[t] = j.rows[ridx[t] * W + TW + c]; }     // sparse update of a live array: normal stores
                        else put(dst + t, ridx[t] >= 0 ? j.rows[ridx[t] * W + TW + c] : src[t]);
                    }
                }
            } else {
                for (int t = 0; t < nb; ++t) {
                    double* dst = j.state_out + (b0 + t) * 14;
                    if (inplace) { if (ridx[t] >= 0) memcpy(dst, j.rows + ridx[t] * W + TW, 14 * sizeof(double)); }
                    else put_n(dst, ridx[t] >= 0 ? j.rows + ridx[t] * W + TW : j.state_in + (b0 + t) * 14, 14);
                }
            }
        }
    }
    store_fence();
}
}  // namespace

// Dense tangent of points [i0, i0 + n) of an N-point array from the chunk's factored rows (mm_host.cu expands every chunk of the dense
// host-buffer call as it arrives): `fl` = the chunk's n flags, `rows` = its n_rows rows of 13 doubles.  Split over `T` host threads.
int expand_factored_chunk(const MMPlastic* p, const uint8_t* fl, const double* rows, int64_t n_rows, double* tan, int64_t N, int64_t i0, int64_t n, int layout, int T) {
    ExpandJob j;
    j.row_kind = MM_ROWS_FACTORED; j.layout = layout; j.flags = fl; j.rows = rows; j.state_in = nullptr; j.tan = tan; j.state_out = nullptr; j.N = N; j.flags_base = i0;
    const ElasticK ek = make_elastic_k(p->E, p->nu);
    j.lam = ek.lam; j.G = ek.G;
    const double zero[6] = {0, 0, 0, 0, 0, 0};
    for (int J = 0; J < 6; ++J)
        for (int I = 0; I < 6; ++I) j.Ee[I + 6 * J] = plastic_tangent_entry_lg(ek.lam, ek.G, 0.0, zero, zero, I, J);
    if (T <= 0) T = host_threads();
    if ((int64_t)T * 16384 > n) T = (int)(n / 16384) + 1;
    std::vector<int64_t> lo(T + 1), row0(T + 1, 0);
    for (int t = 0; t <= T; ++t) lo[t] = i0 + (n * t / T) / BLK * BLK;
    lo[T] = i0 + n;
    for (int t = 0; t < T; ++t) {                      // 1 byte per point: counting is negligible next to the 288-byte scatter
        int64_t c = 0;
        for (int64_t i = lo[t]; i < lo[t + 1]; ++i) c += (fl[i - i0] & MM_FLAG_PLASTIC);
        row0[t + 1] = row0[t] + c;
    }
    if (row0[T] != n_rows) return set_error(MM_ERR_CUDA, "mmh_plastic: chunk at %lld: %lld flags set but %lld rows arrived", (long long)i0, (long long)row0[T], (long long)n_rows);
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back([&, t] { expand_range(j, lo[t], lo[t + 1], row0[t]); });
    expand_range(j, lo[0], lo[1], row0[0]);
    for (auto& x : th) x.join();
    return MM_OK;
}
int expand_host_threads() { return host_threads(); }

// LinearElastic hands every point the ONE stored Eᵉ (src/LinearElastic.jl:59): the host-buffer call does not fetch N copies of it over the
// link, the host cores write them (non-temporal stores) while the strains and stresses are in flight.  Ee = iso_entry, the kernel's values.
void fill_constant_tangent(double lam, double G, double* tan, int64_t N, int layout, int T) {
    double Ee[36];
    for (int J = 0; J < 6; ++J)
        for (int I = 0; I < 6; ++I) Ee[I + 6 * J] = iso_entry(I, J, lam, G);
    if (T <= 0) T = host_threads();
    if ((int64_t)T * 16384 > N) T = (int)(N / 16384) + 1;
    auto work = [&](int64_t lo, int64_t hi) {
        if (layout == MM_LAYOUT_SOA) {
            for (int c = 0; c < 36; ++c) { double* d = tan + (int64_t)c * N; const double v = Ee[c]; for (int64_t i = lo; i < hi; ++i) put(d + i, v); }
        } else {
            for (int64_t i = lo; i < hi; ++i) put_n(tan + i * 36, Ee, 36);
        }
        store_fence();
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(work, N * t / T, N * (t + 1) / T);
    work(0, N / T);
    for (auto& x : th) x.join();
}
}  // namespace mm

using namespace mm;

extern "C" {

int mmh_set_host_threads(int n) {
    const int old = host_threads();
    g_host_threads = n > 0 ? n : 0;
    return old;
}

int mmh_plastic_expand(const MMPlastic* p, int row_kind, const uint8_t* flags, const double* rows, int64_t n_rows, const double* state_in, double* tan,
                       double* state_out, int64_t N, int layout) {
    const int W = mm_plastic_row_width(row_kind);
    if (W < 0) return W;
    const bool with_state = (row_kind == MM_ROWS_TANGENT_STATE || row_kind == MM_ROWS_FACTORED_STATE);
    if (!p || N < 0 || n_rows < 0 || (N > 0 && !flags) || (n_rows > 0 && !rows)) return set_error(MM_ERR_ARG, "mmh_plastic_expand: bad argument");
    if (layout != MM_LAYOUT_SOA && layout != MM_LAYOUT_AOS) return set_error(MM_ERR_ARG, "mmh_plastic_expand: unknown layout %d", layout);
    if (state_out && !with_state) return set_error(MM_ERR_ARG, "mmh_plastic_expand: row kind %d carries no state rows", row_kind);
    if (state_out && !state_in) return set_error(MM_ERR_ARG, "mmh_plastic_expand: state_in required to fill the elastic points of state_out");
    if (N == 0 || (!tan && !state_out)) return MM_OK;
    ExpandJob j;
    j.row_kind = row_kind; j.layout = layout; j.flags = flags; j.rows = rows; j.state_in = state_in; j.tan = tan; j.state_out = state_out; j.N = N; j.flags_base = 0;
    const ElasticK ek = make_elastic_k(p->E, p->nu);
    j.lam = ek.lam; j.G = ek.G;
    const double zero[6] = {0, 0, 0, 0, 0, 0};
    for (int J = 0; J < 6; ++J)
        for (int I = 0; I < 6; ++I) j.Ee[I + 6 * J] = plastic_tangent_entry_lg(ek.lam, ek.G, 0.0, zero, zero, I, J);
    int T = host_threads();
    if ((int64_t)T * 4096 > N) T = (int)(N / 4096) + 1;
    std::vector<int64_t> lo(T + 1), cnt(T, 0);
    for (int t = 0; t <= T; ++t) lo[t] = (N * t / T) / BLK * BLK;
    lo[T] = N;
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t)
            th.emplace_back([&, t] { int64_t c = 0; for (int64_t i = lo[t]; i < lo[t + 1]; ++i) c += (flags[i] & MM_FLAG_PLASTIC); cnt[t] = c; });
        { int64_t c = 0; for (int64_t i = lo[0]; i < lo[1]; ++i) c += (flags[i] & MM_FLAG_PLASTIC); cnt[0] = c; }
        for (auto& x : th) x.join();
    }
    std::vector<int64_t> row0(T + 1, 0);
    for (int t = 0; t < T; ++t) row0[t + 1] = row0[t] + cnt[t];
    if (row0[T] != n_rows) return set_error(MM_ERR_ARG, "mmh_plastic_expand: %lld flags are set but n_rows = %lld", (long long)row0[T], (long long)n_rows);
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back([&, t] { expand_range(j, lo[t], lo[t + 1], row0[t]); });
        expand_range(j, lo[0], lo[1], row0[0]);
        for (auto& x : th) x.join();
    }
    return MM_OK;
}

}  // extern "C"
