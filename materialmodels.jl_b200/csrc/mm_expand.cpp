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
    double lam, G;
    double Ee[36];
};

constexpr int BLK = 256;

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
        for (int t = 0; t < nb; ++t) ridx[t] = (j.flags[b0 + t] & MM_FLAG_PLASTIC) ? row++ : -1;
        if (j.tan) {
            if (soa) {
                for (int J = 0; J < 6; ++J)
                    for (int I = 0; I < 6; ++I) {
                        const int c = I + 6 * J;
                        double* dst = j.tan + (int64_t)c * N + b0;
                        const double e = j.Ee[c];
                        if (factored) {
                            for (int t = 0; t < nb; ++t) {
                                if (ridx[t] < 0) { dst[t] = e; continue; }
                                const double* r = j.rows + ridx[t] * W;
                                dst[t] = plastic_tangent_entry_lg(j.lam, j.G, r[0], r + 1, r + 7, I, J);
                            }
                        } else {
                            for (int t = 0; t < nb; ++t) dst[t] = ridx[t] < 0 ? e : j.rows[ridx[t] * W + c];
                        }
                    }
            } else {
                for (int t = 0; t < nb; ++t) {
                    double* dst = j.tan + (b0 + t) * 36;
                    if (ridx[t] < 0) { memcpy(dst, j.Ee, 36 * sizeof(double)); continue; }
                    const double* r = j.rows + ridx[t] * W;
                    if (factored) {
                        for (int J = 0; J < 6; ++J)
                            for (int I = 0; I < 6; ++I) dst[I + 6 * J] = plastic_tangent_entry_lg(j.lam, j.G, r[0], r + 1, r + 7, I, J);
                    } else {
                        memcpy(dst, r, 36 * sizeof(double));
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
                        if (ridx[t] >= 0) dst[t] = j.rows[ridx[t] * W + TW + c];
                        else if (!inplace) dst[t] = src[t];
                    }
                }
            } else {
                for (int t = 0; t < nb; ++t) {
                    double* dst = j.state_out + (b0 + t) * 14;
                    if (ridx[t] >= 0) memcpy(dst, j.rows + ridx[t] * W + TW, 14 * sizeof(double));
                    else if (!inplace) memcpy(dst, j.state_in + (b0 + t) * 14, 14 * sizeof(double));
                }
            }
        }
    }
}
}  // namespace
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
    j.row_kind = row_kind; j.layout = layout; j.flags = flags; j.rows = rows; j.state_in = state_in; j.tan = tan; j.state_out = state_out; j.N = N;
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
