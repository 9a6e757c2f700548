// mm_cells_* — SURVEY §8(f) rank 4b, the CALLER side of material_response fused around it: the quadrature loop of a
// small-strain solid element (what a Ferrite.jl-style `assemble_cell!` does around the reference's per-point call):
//
//     for q in 1:nq                                              # caller-owned loop in the reference's world
//         ε  = function_symmetric_gradient(cv, q, ue)           # ε = sym(Σ_a u_a ⊗ ∇N_a)
//         σ, D, state_new[q] = material_response(m, ε, state[q]) # <- the reference's hot path (src/MaterialModels.jl:41-52)
//         fe[a] += (σ ⋅ ∇N_a) dΩ                                # Bᵀσ
//         ke[a,b] += δε_a ⊡ D ⊡ Δε_b dΩ                         # BᵀDB  (optional)
//
// One thread per quadrature point; the NQ points of a cell sit in NQ adjacent lanes and sum their Bᵀσ contributions through
// shared memory, so σ never goes to memory and the 36-entry tangent is written only if the caller asks for it
// (implicit codes), which removes 288 of the 608 algorithmic bytes per point of the unfused J2 update.
// Arrays (all FP64, component-major like MM_LAYOUT_SOA):
//   per point  (P = ncell*NQ points, point index = cell*NQ + q):  dNdx[(3a+k)*P + i], dOmega[i], state[c*P + i], sig/tan
//   per cell   (ncell cells):                                     ue[(3a+k)*ncell + cell], fe[(3a+k)*ncell + cell]
#include <cuda_runtime.h>
#include "../../include/mmb200.h"
#include "mm_kernels.cuh"

namespace mm {

// accessor for the inner 3D model: strain from registers, state straight from / to global memory (coalesced), σ kept in registers
// KEEP_D: the 36-entry tangent also stays in registers (for the element stiffness BᵀDB)
template <bool KEEP_D> struct CellAcc {
    double eps[6], sig[6];
    double D[KEEP_D ? 36 : 1];
    const double* st_in; double* st_out; double* tan; double* sig_out;
    int64_t P, i;
    template <int A> __device__ __forceinline__ double in(int c) const { return A == 0 ? eps[c] : __ldcs(st_in + (int64_t)c * P + i); }
    template <int A> __device__ __forceinline__ void out(int c, double v) {
        if (A == 0) { sig[c] = v; if (sig_out) __stcs(sig_out + (int64_t)c * P + i, v); }
        else if (A == 1) { if (KEEP_D) D[KEEP_D ? c : 0] = v; if (tan) __stcs(tan + (int64_t)c * P + i, v); }
        else __stcs(st_out + (int64_t)c * P + i, v);
    }
    template <int A> __device__ __forceinline__ bool has_out() const { return A == 1 ? (KEEP_D || tan != nullptr) : true; }
};

// finite strain: F in registers, Pᵀ kept in registers (and optionally stored), ∂Pᵀ∂F straight to global memory on request
struct HyperCellAcc {
    double F[9], Pt[9];
    double* pt_out; double* dP_out;
    int64_t P, i;
    template <int A> __device__ __forceinline__ double in(int c) const { return F[c]; }
    template <int A> __device__ __forceinline__ void out(int c, double v) {
        if (A == 0) { Pt[c] = v; if (pt_out) __stcs(pt_out + (int64_t)c * P + i, v); }
        else __stcs(dP_out + (int64_t)c * P + i, v);
    }
    template <int A> __device__ __forceinline__ bool has_out() const { return A == 1 ? dP_out != nullptr : true; }
};

struct CellArgs {
    const double* dNdx; const double* dOmega; const double* ue; const double* state_in;
    double* fe; double* ke; double* sig; double* tan; double* state_out;
    uint8_t* flags; uint16_t* iters; int32_t* n_fail;
    int64_t ncell;
};

// MODEL 0: LinearElastic (stateless), 1: Plastic (14 state doubles), 2: TransverselyIsotropic (state = direction a3, read-only)
// ∇N is parked in shared memory across the material update (3·NN doubles per thread would otherwise cost 2·3·NN registers on
// top of the J2 solve: 255 registers and 15.2 ms for the hexahedron, against 168 and 10.7 ms this way); for NQ > 1 the Bᵀσ
// contributions of the NQ lanes of a cell are summed through a padded shared-memory image: lane q adds up and stores components q, q+NQ, … (3·NN/NQ each) — one
// STS + 3·NN/NQ·NQ LDS per thread instead of 3·NN·log2(NQ) 64-bit shuffles.
#ifndef MM_CELLS_KE_UNROLL
#define MM_CELLS_KE_UNROLL 1
#endif
#ifndef MM_CELLS_MINBLOCKS
#define MM_CELLS_MINBLOCKS 3
#endif
#ifndef MM_CELLS_KE_MINBLOCKS
#define MM_CELLS_KE_MINBLOCKS 2      // element-stiffness instantiations: resident blocks of 128 threads (register cap)
#endif
#ifndef MM_CELLS_KE_CH
#define MM_CELLS_KE_CH 8             // node columns accumulated at once per lane (3 x 3·CH accumulators)
#endif
template <int MODEL, int NN, int NQ, int BS, bool WITH_KE>
__global__ void __launch_bounds__(BS, WITH_KE ? MM_CELLS_KE_MINBLOCKS : MM_CELLS_MINBLOCKS) cells_kernel(const ElasticK ke, const PlasticK kp, const TransvIsoK kt, const CellArgs a) {
    static_assert(NQ == 1 || NQ == 2 || NQ == 4 || NQ == 8, "quadrature points per cell must be a power of two <= 8");
    static_assert(BS % 32 == 0 && 32 % NQ == 0, "a cell's points must share a warp");
    constexpr int NC = 3 * NN;
    extern __shared__ double sm[];
    double* sm_dN = sm;                                  // [NC][BS]
    double* sm_f = sm + NC * BS;                         // NQ > 1: [NC][BS + 1] fe image (padded: the NQ lanes of a cell read NQ different rows); the [BS][38] tangent image follows
    const int64_t P = a.ncell * NQ;
    const int64_t i = (int64_t)blockIdx.x * BS + threadIdx.x;
    const unsigned mask = __ballot_sync(0xffffffffu, i < P);   // cells are whole and warp-local: the survivors sync among themselves below
    if (i >= P) return;
    const int64_t cell = i / NQ;
    const int q = (int)(i - cell * NQ);
    const int t = threadIdx.x;
    CellAcc<WITH_KE> acc;
#pragma unroll
    for (int c = 0; c < 6; ++c) acc.eps[c] = 0.0;
#pragma unroll
    for (int n = 0; n < NN; ++n) {
        double u[3], d[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            d[k] = __ldcs(a.dNdx + (int64_t)(3 * n + k) * P + i);
            u[k] = __ldg(a.ue + (int64_t)(3 * n + k) * a.ncell + cell);       // shared by the NQ lanes of the cell
            sm_dN[(3 * n + k) * BS + t] = d[k];
        }
        // ε = sym(Σ_a u_a ⊗ ∇N_a), storage (11,21,31,22,32,33)
        acc.eps[0] += u[0] * d[0];
        acc.eps[3] += u[1] * d[1];
        acc.eps[5] += u[2] * d[2];
        acc.eps[1] += 0.5 * (u[1] * d[0] + u[0] * d[1]);
        acc.eps[2] += 0.5 * (u[2] * d[0] + u[0] * d[2]);
        acc.eps[4] += 0.5 * (u[2] * d[1] + u[1] * d[2]);
    }
    const double dO = __ldcs(a.dOmega + i);
    acc.st_in = a.state_in; acc.st_out = a.state_out; acc.tan = a.tan; acc.sig_out = a.sig; acc.P = P; acc.i = i;
    if (MODEL == 0) {
        linear_elastic_point(ke, acc);
    } else if (MODEL == 2) {
        transviso_point(kt, acc);
    } else {
        int it = 0;
        const int flag = plastic_point(kp, acc, &it);
        if (a.flags) a.flags[i] = (uint8_t)flag;
        if (a.iters) a.iters[i] = (uint16_t)(it > 65535 ? 65535 : it);
        if ((flag & MM_FLAG_FAILED) && a.n_fail) atomicAdd(a.n_fail, 1);
    }
    // fe[a][k] = Σ_q (σ ⋅ ∇N_a)_k dΩ
    const double s00 = acc.sig[0] * dO, s10 = acc.sig[1] * dO, s20 = acc.sig[2] * dO, s11 = acc.sig[3] * dO, s21 = acc.sig[4] * dO, s22 = acc.sig[5] * dO;
#pragma unroll
    for (int n = 0; n < NN; ++n) {
        double d[3], f[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) d[k] = sm_dN[(3 * n + k) * BS + t];
        f[0] = s00 * d[0] + s10 * d[1] + s20 * d[2];
        f[1] = s10 * d[0] + s11 * d[1] + s21 * d[2];
        f[2] = s20 * d[0] + s21 * d[1] + s22 * d[2];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (NQ == 1) __stcs(a.fe + (int64_t)(3 * n + k) * a.ncell + cell, f[k]);
            else sm_f[(3 * n + k) * (BS + 1) + t] = f[k];
        }
    }
    const int t0 = t - q;                                  // first lane of this cell
    if (NQ > 1) {
        __syncwarp(mask);
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += NQ) {
            const int c = c0 + q;
            if (c < NC) {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < NQ; ++j) v += sm_f[c * (BS + 1) + t0 + j];
                __stcs(a.fe + (int64_t)c * a.ncell + cell, v);
            }
        }
    }
    if (WITH_KE) {
        // ke[3a+i][3b+j] = Σ_q Σ_nl ∇N_a[n] D_(in)(jl) ∇N_b[l] dΩ   (δε_i ⊡ D ⊡ Δε_j with the minor symmetries of D):
        // T_a[i][J] = Σ_n ∇N_a[n] D[(in), J] (54 FMA), then block[i][j] = Σ_l T_a[i][(jl)] ∇N_b[l] (27 FMA per node pair).
        // Stored component-major over cells: ke[((3a+i)*NC + 3b+j)*ncell + cell].
        if (NQ == 1) {                                     // thread = cell: block rows straight to global memory
#pragma unroll 1
            for (int na = 0; na < NN; ++na) {
                double da[3], T[3][6];
#pragma unroll
                for (int k = 0; k < 3; ++k) da[k] = sm_dN[(3 * na + k) * BS + t] * dO;
#pragma unroll
                for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                    for (int J = 0; J < 6; ++J)
                        T[ii][J] = da[0] * acc.D[WITH_KE ? sym_slot(ii, 0) + 6 * J : 0] + da[1] * acc.D[WITH_KE ? sym_slot(ii, 1) + 6 * J : 0] +
                                   da[2] * acc.D[WITH_KE ? sym_slot(ii, 2) + 6 * J : 0];
#pragma unroll 1
                for (int nb = 0; nb < NN; ++nb) {
                    double db[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) db[k] = sm_dN[(3 * nb + k) * BS + t];
#pragma unroll
                    for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                        for (int jj = 0; jj < 3; ++jj)
                            __stcs(a.ke + ((int64_t)(3 * na + ii) * NC + (3 * nb + jj)) * a.ncell + cell,
                                   T[ii][sym_slot(jj, 0)] * db[0] + T[ii][sym_slot(jj, 1)] * db[1] + T[ii][sym_slot(jj, 2)] * db[2]);
                }
            }
        } else {
            // NQ lanes per cell: lane q owns the block rows of nodes a = q, q+NQ, … and accumulates them over the cell's NQ points in
            // registers — no cross-lane reduction.  Every point parks its tangent and dΩ in shared memory ([thread][38], read back as
            // broadcast 128-bit loads by the NQ lanes of the cell); ∇N is already there.  Columns in chunks of <= 8 nodes (72 accumulators).
            constexpr int PD = 38, CH = (NN < MM_CELLS_KE_CH) ? NN : MM_CELLS_KE_CH, MM_CELLS_KE_UNROLL_V = MM_CELLS_KE_UNROLL;
            double* sm_D = sm_f + ((NC * (BS + 1) + 1) & ~1);   // its own region (the fe image is column-per-thread, this one row-per-thread:
                                                                // sharing the buffer would need a block-wide barrier); rounded up to 16 bytes
#pragma unroll
            for (int c = 0; c < 36; c += 2) *reinterpret_cast<double2*>(sm_D + t * PD + c) = make_double2(acc.D[WITH_KE ? c : 0], acc.D[WITH_KE ? c + 1 : 0]);
            sm_D[t * PD + 36] = dO;
            __syncwarp(mask);
#pragma unroll 1
            for (int na = q; na < NN; na += NQ) {
#pragma unroll
                for (int b0 = 0; b0 < NN; b0 += CH) {
                    double K[3][3 * CH];
#pragma unroll
                    for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                        for (int c = 0; c < 3 * CH; ++c) K[ii][c] = 0.0;
#pragma unroll MM_CELLS_KE_UNROLL_V
                    for (int pq = 0; pq < NQ; ++pq) {
                        const double* Dp = sm_D + (t0 + pq) * PD;
                        const double w = Dp[36];
                        double da[3], T[3][6];
#pragma unroll
                        for (int k = 0; k < 3; ++k) da[k] = sm_dN[(3 * na + k) * BS + t0 + pq] * w;
#pragma unroll
                        for (int J = 0; J < 6; ++J) {
                            const double2 d01 = *reinterpret_cast<const double2*>(Dp + 6 * J), d23 = *reinterpret_cast<const double2*>(Dp + 6 * J + 2),
                                          d45 = *reinterpret_cast<const double2*>(Dp + 6 * J + 4);
                            const double Dc[6] = {d01.x, d01.y, d23.x, d23.y, d45.x, d45.y};          // D[I + 6J], I = 0..5
#pragma unroll
                            for (int ii = 0; ii < 3; ++ii) T[ii][J] = da[0] * Dc[sym_slot(ii, 0)] + da[1] * Dc[sym_slot(ii, 1)] + da[2] * Dc[sym_slot(ii, 2)];
                        }
#pragma unroll
                        for (int nb = 0; nb < CH; ++nb) {
                            if (b0 + nb < NN) {
                                double db[3];
#pragma unroll
                                for (int k = 0; k < 3; ++k) db[k] = sm_dN[(3 * (b0 + nb) + k) * BS + t0 + pq];
#pragma unroll
                                for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                                    for (int jj = 0; jj < 3; ++jj)
                                        K[ii][3 * nb + jj] += T[ii][sym_slot(jj, 0)] * db[0] + T[ii][sym_slot(jj, 1)] * db[1] + T[ii][sym_slot(jj, 2)] * db[2];
                            }
                        }
                    }
#pragma unroll
                    for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                        for (int c = 0; c < 3 * CH; ++c)
                            if (3 * b0 + c < NC) __stcs(a.ke + ((int64_t)(3 * na + ii) * NC + (3 * b0 + c)) * a.ncell + cell, K[ii][c]);
                }
            }
        }
    }
}

template <int MODEL, int NN, int NQ, bool WITH_KE>
int launch_cells_k(const ElasticK& ke, const PlasticK& kp, const TransvIsoK& kt, const CellArgs& a, cudaStream_t st) {
    constexpr int BS = 128;
    constexpr size_t fe_img = (NQ > 1) ? (size_t)3 * NN * (BS + 1) : 0, d_img = (NQ > 1 && WITH_KE) ? (size_t)38 * BS + 2 : 0;
    constexpr size_t smem = ((size_t)3 * NN * BS + fe_img + d_img) * sizeof(double);
    const int64_t P = a.ncell * NQ;
    if (P == 0) return MM_OK;
    const int64_t blocks = (P + BS - 1) / BS;
    if (blocks > 2147483647LL) return set_error(MM_ERR_ARG, "mm_cells: too many points for one launch");
    static SmemOptIn optin;
    if (int rc = ensure_smem(optin, cells_kernel<MODEL, NN, NQ, BS, WITH_KE>, smem, "mm_cells")) return rc;
    cells_kernel<MODEL, NN, NQ, BS, WITH_KE><<<(unsigned)blocks, BS, smem, st>>>(ke, kp, kt, a);
    return check_launch("mm_cells");
}
template <int MODEL, int NN, int NQ>
int launch_cells(const ElasticK& ke, const PlasticK& kp, const TransvIsoK& kt, const CellArgs& a, cudaStream_t st) {
    return a.ke ? launch_cells_k<MODEL, NN, NQ, true>(ke, kp, kt, a, st) : launch_cells_k<MODEL, NN, NQ, false>(ke, kp, kt, a, st);
}

// Total-Lagrangian counterpart: F = I + Σ_a u_a ⊗ ∇N_a (∇N w.r.t. the reference configuration), (Pᵀ, ∂Pᵀ∂F) from the hyperelastic
// model through the trait route of src/traits.jl:67-77, fe[a][i] = Σ_j P_ij ∇N_a[j] dΩ.  Same thread/lane mapping and shared-memory
// reduction as cells_kernel.
struct HyperCellArgs {
    const double* dNdx; const double* dOmega; const double* ue;
    double* fe; double* Pt; double* dPdF;
    int64_t ncell;
};
#ifndef MM_HCELLS_MINBLOCKS
#define MM_HCELLS_MINBLOCKS 3
#endif
template <int HMODEL, int NN, int NQ, int BS>
__global__ void __launch_bounds__(BS, MM_HCELLS_MINBLOCKS) hyper_cells_kernel(const HyperK k, const HyperCellArgs a) {
    static_assert(NQ == 1 || NQ == 2 || NQ == 4 || NQ == 8, "quadrature points per cell must be a power of two <= 8");
    constexpr int NC = 3 * NN;
    extern __shared__ double sm[];
    double* sm_dN = sm;
    double* sm_f = sm + NC * BS;
    const int64_t P = a.ncell * NQ;
    const int64_t i = (int64_t)blockIdx.x * BS + threadIdx.x;
    const unsigned mask = __ballot_sync(0xffffffffu, i < P);
    if (i >= P) return;
    const int64_t cell = i / NQ;
    const int q = (int)(i - cell * NQ);
    const int t = threadIdx.x;
    HyperCellAcc acc;
#pragma unroll
    for (int c = 0; c < 9; ++c) acc.F[c] = (c == 0 || c == 4 || c == 8) ? 1.0 : 0.0;
#pragma unroll
    for (int n = 0; n < NN; ++n) {
        double u[3], d[3];
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
            d[kk] = __ldcs(a.dNdx + (int64_t)(3 * n + kk) * P + i);
            u[kk] = __ldg(a.ue + (int64_t)(3 * n + kk) * a.ncell + cell);
            sm_dN[(3 * n + kk) * BS + t] = d[kk];
        }
#pragma unroll
        for (int jj = 0; jj < 3; ++jj)
#pragma unroll
            for (int ii = 0; ii < 3; ++ii) acc.F[ii + 3 * jj] += u[ii] * d[jj];             // F_ij += u_a[i] ∇N_a[j]
    }
    const double dO = __ldcs(a.dOmega + i);
    acc.pt_out = a.Pt; acc.dP_out = a.dPdF; acc.P = P; acc.i = i;
    hyper_point_ex<HMODEL, 1, 2>(k, acc);
#pragma unroll
    for (int n = 0; n < NN; ++n) {
        double d[3];
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) d[kk] = sm_dN[(3 * n + kk) * BS + t] * dO;
#pragma unroll
        for (int ii = 0; ii < 3; ++ii) {
            const double f = acc.Pt[0 + 3 * ii] * d[0] + acc.Pt[1 + 3 * ii] * d[1] + acc.Pt[2 + 3 * ii] * d[2];   // P_ij = Pᵀ_ji = Pt[j + 3 i]
            if (NQ == 1) __stcs(a.fe + (int64_t)(3 * n + ii) * a.ncell + cell, f);
            else sm_f[(3 * n + ii) * (BS + 1) + t] = f;
        }
    }
    if (NQ > 1) {
        __syncwarp(mask);
        const int t0 = t - q;
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += NQ) {
            const int c = c0 + q;
            if (c < NC) {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < NQ; ++j) v += sm_f[c * (BS + 1) + t0 + j];
                __stcs(a.fe + (int64_t)c * a.ncell + cell, v);
            }
        }
    }
}
template <int HMODEL, int NN, int NQ>
int launch_hyper_cells(const HyperK& k, const HyperCellArgs& a, cudaStream_t st) {
    constexpr int BS = 128;
    constexpr size_t smem = (size_t)3 * NN * ((NQ > 1) ? (2 * BS + 1) : BS) * sizeof(double);
    const int64_t P = a.ncell * NQ;
    if (P == 0) return MM_OK;
    const int64_t blocks = (P + BS - 1) / BS;
    if (blocks > 2147483647LL) return set_error(MM_ERR_ARG, "mm_cells_hyper: too many points for one launch");
    static SmemOptIn optin;
    if (int rc = ensure_smem(optin, hyper_cells_kernel<HMODEL, NN, NQ, BS>, smem, "mm_cells_hyper")) return rc;
    hyper_cells_kernel<HMODEL, NN, NQ, BS><<<(unsigned)blocks, BS, smem, st>>>(k, a);
    return check_launch("mm_cells_hyper");
}
template <int HMODEL>
int dispatch_hyper_cells(int nn, int nq, const HyperK& k, const HyperCellArgs& a, cudaStream_t st) {
    if (nn == 4 && nq == 1) return launch_hyper_cells<HMODEL, 4, 1>(k, a, st);
    if (nn == 4 && nq == 4) return launch_hyper_cells<HMODEL, 4, 4>(k, a, st);
    if (nn == 8 && nq == 8) return launch_hyper_cells<HMODEL, 8, 8>(k, a, st);
    if (nn == 8 && nq == 1) return launch_hyper_cells<HMODEL, 8, 1>(k, a, st);
    if (nn == 10 && nq == 4) return launch_hyper_cells<HMODEL, 10, 4>(k, a, st);
    return set_error(MM_ERR_ARG, "mm_cells_hyper: unsupported element (nodes %d, quadrature points %d); supported: (4,1) (4,4) (8,1) (8,8) (10,4)", nn, nq);
}

template <int MODEL>
int dispatch_cells(int nn, int nq, const ElasticK& ke, const PlasticK& kp, const TransvIsoK& kt, const CellArgs& a, cudaStream_t st) {
    if (nn == 4 && nq == 1) return launch_cells<MODEL, 4, 1>(ke, kp, kt, a, st);       // linear tetrahedron, 1-point rule
    if (nn == 4 && nq == 4) return launch_cells<MODEL, 4, 4>(ke, kp, kt, a, st);       // linear tetrahedron, 4-point rule
    if (nn == 8 && nq == 8) return launch_cells<MODEL, 8, 8>(ke, kp, kt, a, st);       // trilinear hexahedron, 2x2x2 Gauss
    if (nn == 8 && nq == 1) return launch_cells<MODEL, 8, 1>(ke, kp, kt, a, st);       // trilinear hexahedron, reduced integration
    if (nn == 10 && nq == 4) return launch_cells<MODEL, 10, 4>(ke, kp, kt, a, st);     // quadratic tetrahedron, 4-point rule
    return set_error(MM_ERR_ARG, "mm_cells: unsupported element (nodes %d, quadrature points %d); supported: (4,1) (4,4) (8,1) (8,8) (10,4)", nn, nq);
}

}  // namespace mm

using namespace mm;

extern "C" {

int mm_cells_linear_elastic(const MMLinearElastic* p, int nn, int nq, const double* dNdx, const double* dOmega, const double* ue, double* fe, double* ke_out,
                            double* sig, double* tan, int64_t ncell, void* stream) {
    if (!p || ncell < 0 || (ncell > 0 && (!dNdx || !dOmega || !ue || !fe))) return set_error(MM_ERR_ARG, "mm_cells_linear_elastic: bad argument");
    CellArgs a{dNdx, dOmega, ue, nullptr, fe, ke_out, sig, tan, nullptr, nullptr, nullptr, nullptr, ncell};
    PlasticK kp{};
    TransvIsoK kt{};
    return dispatch_cells<0>(nn, nq, make_elastic_k(p->E, p->nu), kp, kt, a, (cudaStream_t)stream);
}

int mm_cells_plastic(const MMPlastic* p, int nn, int nq, const double* dNdx, const double* dOmega, const double* ue, const double* state_in, double* fe,
                     double* ke_out, double* sig, double* tan, double* state_out, uint8_t* flags, uint16_t* iters, int32_t* n_fail, const MMNewtonOpts* opts, int64_t ncell,
                     void* stream) {
    if (!p || ncell < 0 || (ncell > 0 && (!dNdx || !dOmega || !ue || !state_in || !fe || !state_out))) return set_error(MM_ERR_ARG, "mm_cells_plastic: bad argument");
    CellArgs a{dNdx, dOmega, ue, state_in, fe, ke_out, sig, tan, state_out, flags, iters, n_fail, ncell};
    ElasticK ke = make_elastic_k(p->E, p->nu);
    PlasticK kp = make_plastic_k(p->E, p->nu, p->sigma_y, p->H, p->r, p->kappa_inf, p->alpha_inf, opts ? opts->tol : 1e-8, opts ? opts->max_iter : 1000);
    TransvIsoK kt{};
    return dispatch_cells<1>(nn, nq, ke, kp, kt, a, (cudaStream_t)stream);
}

int mm_cells_transversely_isotropic(const MMTransverselyIsotropic* p, int nn, int nq, const double* dNdx, const double* dOmega, const double* ue,
                                    const double* a3, double* fe, double* ke_out, double* sig, double* tan, int64_t ncell, void* stream) {
    if (!p || ncell < 0 || (ncell > 0 && (!dNdx || !dOmega || !ue || !a3 || !fe))) return set_error(MM_ERR_ARG, "mm_cells_transversely_isotropic: bad argument");
    CellArgs a{dNdx, dOmega, ue, a3, fe, ke_out, sig, tan, nullptr, nullptr, nullptr, nullptr, ncell};
    ElasticK ke{};
    PlasticK kp{};
    TransvIsoK kt{p->L_perp, p->L_par, p->M_par, p->G_perp, p->G_par};
    return dispatch_cells<2>(nn, nq, ke, kp, kt, a, (cudaStream_t)stream);
}

int mm_cells_hyper(int model, const MMHyper* p, int nn, int nq, const double* dNdx, const double* dOmega, const double* ue, double* fe, double* Pt,
                   double* dPdF, int64_t ncell, void* stream) {
    if (!p || ncell < 0 || (ncell > 0 && (!dNdx || !dOmega || !ue || !fe))) return set_error(MM_ERR_ARG, "mm_cells_hyper: bad argument");
    HyperK k; k.lambda = p->lambda; k.mu = p->mu; k.c2 = p->c2; k.c3 = p->c3;
    HyperCellArgs a{dNdx, dOmega, ue, fe, Pt, dPdF, ncell};
    switch (model) {
        case MM_NEOHOOK: return dispatch_hyper_cells<0>(nn, nq, k, a, (cudaStream_t)stream);
        case MM_YEOH: return dispatch_hyper_cells<1>(nn, nq, k, a, (cudaStream_t)stream);
        case MM_STVENANT: return dispatch_hyper_cells<2>(nn, nq, k, a, (cudaStream_t)stream);
    }
    return set_error(MM_ERR_ARG, "mm_cells_hyper: unknown model %d", model);
}

}  // extern "C"
