// Compact return of the J2 update — reference src/Plastic.jl:134-135: an elastic point returns the ONE stored m.Eᵉ and the
// UNCHANGED state object; only a point that took the Newton branch (:137-159) gets a tangent and a state of its own.  The dense
// entry point (mm_plastic) materialises 36 + 14 doubles for every point all the same — 400 of its 608 bytes per point, which is what
// bounds a host-resident caller over PCIe.  Here every point gets σ, its flag byte and its Newton count; the yielded points, and
// only those, get one ROW each, in point order:
//     MM_ROWS_TANGENT_STATE   [∂σ∂ε(36) | state(14)]           MM_ROWS_TANGENT    [∂σ∂ε(36)]
//     MM_ROWS_FACTORED_STATE  [gξ, u(6), n(6) | state(14)]      MM_ROWS_FACTORED   [gξ, u(6), n(6)]
// with ∂σ∂ε = Eᵉ − gξ I_dev + u⊗n — 13 numbers carry the 36-entry tangent exactly (plastic_tangent_entry_lg is the one
// formula the dense kernel, this file and the host expander evaluate, fma for fma).
//
// Three small launches after the update kernel: per-block counts of the yield flags, a single-block exclusive scan of the counts, and a
// scatter in which a thread's row index = block offset + (ballot-popc rank inside the warp) + (warp offset).  Order is the point order,
// so the result is deterministic and chunking the batch (mm_host.cu) concatenates.
#include <cuda_runtime.h>
#include "../../include/mmb200.h"
#include "mm_kernels.cuh"

namespace mm {

int launch_plastic_parts(const MMPlastic* p, const MMNewtonOpts* opts, const double* eps, const double* state_in, double* sig, double* parts, double* state_out,
                         uint8_t* flags, uint16_t* iters, int32_t* n_fail, int64_t N, int layout, cudaStream_t st);   // mm_ops.cu

namespace {
constexpr int CB = 1024;     // points per compaction block

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

__global__ void __launch_bounds__(CB) count_yielded_kernel(const uint8_t* __restrict__ flags, const int64_t N, int32_t* __restrict__ block_counts) {
    const int64_t i = (int64_t)blockIdx.x * CB + threadIdx.x;
    const int pred = (i < N) && (flags[i] & MM_FLAG_PLASTIC);
    const int c = __syncthreads_count(pred);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// exclusive scan of nb block counts by ONE block (nb <= ~10^5 for 10^8 points: a few hundred microseconds at worst, tens for a pipeline chunk)
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int32_t* __restrict__ counts, int64_t* __restrict__ offsets, const int64_t nb, int64_t* __restrict__ total_out) {
    __shared__ int64_t warp_sum[32];
    __shared__ int64_t tile_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t carry = 0;
    for (int64_t base = 0; base < nb; base += 1024) {
        const int64_t idx = base + threadIdx.x;
        const int64_t v = idx < nb ? (int64_t)counts[idx] : 0;
        int64_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int64_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            const int64_t w = warp_sum[lane];
            int64_t z = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int64_t y = __shfl_up_sync(0xffffffffu, z, o); if (lane >= o) z += y; }
            warp_sum[lane] = z - w;
            if (lane == 31) tile_total = z;
        }
        __syncthreads();
        if (idx < nb) offsets[idx] = carry + warp_sum[warp] + (x - v);
        carry += tile_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

// FACTORED: tangent part of a row is the 13 stored numbers, else the 36 entries; WITH_STATE: 14 more doubles
template <bool FACTORED, bool WITH_STATE>
__global__ void __launch_bounds__(CB) scatter_rows_kernel(const uint8_t* __restrict__ flags, const double* __restrict__ parts, const double* __restrict__ state,
                                                          const int64_t* __restrict__ offsets, double* __restrict__ rows, const int64_t cap, const int64_t N,
                                                          const int layout, const double lam, const double G) {
    constexpr int TW = FACTORED ? 13 : 36, W = TW + (WITH_STATE ? 14 : 0);
    __shared__ int warp_cnt[CB / 32];
    const int64_t i = (int64_t)blockIdx.x * CB + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool pred = (i < N) && (flags[i] & MM_FLAG_PLASTIC);
    const unsigned bal = __ballot_sync(0xffffffffu, pred);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    if (!pred) return;
    int before = __popc(bal & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) before += warp_cnt[w];
    const int64_t row = offsets[blockIdx.x] + before;
    if (row >= cap) return;                 // the caller sees n_rows > capacity and retries with a larger buffer
    double* r = rows + row * W;
    double pr[13];
#pragma unroll
    for (int c = 0; c < 13; ++c) pr[c] = layout == MM_LAYOUT_SOA ? __ldcs(parts + ((int64_t)c * N + i)) : __ldcs(parts + (i * 36 + c));
    if (FACTORED) {
#pragma unroll
        for (int c = 0; c < 13; ++c) r[c] = pr[c];
    } else {
#pragma unroll
        for (int J = 0; J < 6; ++J)
#pragma unroll
            for (int I = 0; I < 6; ++I) r[I + 6 * J] = plastic_tangent_entry_lg(lam, G, pr[0], pr + 1, pr + 7, I, J);
    }
    if (WITH_STATE) {
#pragma unroll
        for (int c = 0; c < 14; ++c) r[TW + c] = layout == MM_LAYOUT_SOA ? __ldcs(state + ((int64_t)c * N + i)) : __ldcs(state + (i * 14 + c));
    }
}
}  // namespace
}  // namespace mm

using namespace mm;

extern "C" {

int mm_plastic_row_width(int row_kind) {
    switch (row_kind) {
        case MM_ROWS_TANGENT_STATE: return 50;
        case MM_ROWS_FACTORED_STATE: return 27;
        case MM_ROWS_TANGENT: return 36;
        case MM_ROWS_FACTORED: return 13;
    }
    return set_error(MM_ERR_ARG, "mm_plastic_row_width: unknown row kind %d", row_kind);
}

// the factored-tangent scratch is sized for either layout: 13 doubles per point component-major (SoA) or the tile kernel's 36-double
// stride (AoS, see launch_plastic_parts)
uint64_t mm_plastic_compact_workspace(int64_t N, int want_state_out) {
    if (N <= 0) return 256;
    const int64_t nb = (N + CB - 1) / CB;
    size_t b = al256((size_t)N * 36 * 8) + al256((size_t)nb * 4) + al256((size_t)nb * 8);
    if (!want_state_out) b += al256((size_t)N * 14 * 8);
    return (uint64_t)b;
}

int mm_plastic_compact(const MMPlastic* p, const double* eps, const double* state_in, double* sig, uint8_t* flags, uint16_t* iters, int row_kind,
                       double* rows, int64_t rows_capacity, int64_t* n_rows, double* state_out, int32_t* n_fail, const MMNewtonOpts* opts,
                       void* workspace, uint64_t workspace_bytes, int64_t N, int layout, void* stream) {
    const int W = mm_plastic_row_width(row_kind);
    if (W < 0) return W;
    if (!p || N < 0 || rows_capacity < 0 || !n_rows || (N > 0 && (!eps || !state_in || !sig || !flags || !workspace || (rows_capacity > 0 && !rows))))
        return set_error(MM_ERR_ARG, "mm_plastic_compact: bad argument");
    if (layout != MM_LAYOUT_SOA && layout != MM_LAYOUT_AOS) return set_error(MM_ERR_ARG, "mm_plastic_compact: unknown layout %d", layout);
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {
        if (cudaMemsetAsync(n_rows, 0, sizeof(int64_t), st) != cudaSuccess) return set_error(MM_ERR_CUDA, "mm_plastic_compact: memset failed");
        return MM_OK;
    }
    if (workspace_bytes < mm_plastic_compact_workspace(N, state_out != nullptr))
        return set_error(MM_ERR_ARG, "mm_plastic_compact: workspace too small (%llu < %llu bytes)", (unsigned long long)workspace_bytes,
                         (unsigned long long)mm_plastic_compact_workspace(N, state_out != nullptr));
    const int64_t nb = (N + CB - 1) / CB;
    if (nb > 2147483647LL) return set_error(MM_ERR_ARG, "mm_plastic_compact: N too large for one launch");
    char* w = (char*)workspace;
    double* parts = (double*)w; w += al256((size_t)N * 36 * 8);
    int32_t* counts = (int32_t*)w; w += al256((size_t)nb * 4);
    int64_t* offsets = (int64_t*)w; w += al256((size_t)nb * 8);
    double* state_dense = state_out ? state_out : (double*)w;
    if (int rc = launch_plastic_parts(p, opts, eps, state_in, sig, parts, state_dense, flags, iters, n_fail, N, layout, st)) return rc;
    count_yielded_kernel<<<(unsigned)nb, CB, 0, st>>>(flags, N, counts);
    scan_counts_kernel<<<1, 1024, 0, st>>>(counts, offsets, nb, n_rows);
    if (rows_capacity > 0) {
        const ElasticK ek = make_elastic_k(p->E, p->nu);
        switch (row_kind) {
            case MM_ROWS_TANGENT_STATE: scatter_rows_kernel<false, true><<<(unsigned)nb, CB, 0, st>>>(flags, parts, state_dense, offsets, rows, rows_capacity, N, layout, ek.lam, ek.G); break;
            case MM_ROWS_FACTORED_STATE: scatter_rows_kernel<true, true><<<(unsigned)nb, CB, 0, st>>>(flags, parts, state_dense, offsets, rows, rows_capacity, N, layout, ek.lam, ek.G); break;
            case MM_ROWS_TANGENT: scatter_rows_kernel<false, false><<<(unsigned)nb, CB, 0, st>>>(flags, parts, state_dense, offsets, rows, rows_capacity, N, layout, ek.lam, ek.G); break;
            default: scatter_rows_kernel<true, false><<<(unsigned)nb, CB, 0, st>>>(flags, parts, state_dense, offsets, rows, rows_capacity, N, layout, ek.lam, ek.G); break;
        }
    }
    return check_launch("mm_plastic_compact");
}

}  // extern "C"
