// mm_crystal — device-pointer entry point of CrystalViscoPlastic{12} (FP64-compute-bound model).
#include <cuda_runtime.h>
#include <stdlib.h>
#include "../../include/mmb200.h"
#include "mm_kernels.cuh"
#include "mm_crystal_math.cuh"

namespace mm {

// MODE 0: general path (any H; 12x12 pivoted LU in thread-local memory).
// MODE 1 / 2: register-resident structured path for H = a I + b 11ᵀ without / with the cross-hardening term
// (b = H_iso q).  A point whose SPD factorisation or D-guard fails is only MARKED (bit 14 of its mask); a second,
// tiny launch (crystal_fixup_kernel) re-runs exactly those points through the general pivoted path, so the main
// kernel carries no fallback code (registers) and a rare fallback does not stall its warp.
#define MM_ACTIVE_NEEDS_GENERAL 0x4000u

template <int MODE> struct CrystalOp {
    static constexpr int NIN = 2, NOUT = 3, SCRATCH = (MODE == 0) ? 0 : MM_CRYSTAL_SCRATCH;
    __host__ __device__ static constexpr int in_nc(int a) { return a == 0 ? 6 : 42; }
    __host__ __device__ static constexpr int out_nc(int a) { return a == 0 ? 6 : (a == 1 ? 36 : 42); }
    CrystalK k;
    uint16_t* active;     // never NULL for MODE 1/2 (mm_crystal substitutes a temporary)
    uint8_t* iters;
    int32_t* n_fail;
    template <class Acc> __device__ __forceinline__ void point(Acc& acc, int64_t i, const Scratch& scr) const {
        int it = 0;
        unsigned mask;
        if (MODE == 0) {
            mask = crystal_point(k, acc, &it);
        } else {
            bool fallback = false;
            mask = crystal_point_fast<MODE == 2>(k, acc, scr, &it, &fallback);
            if (fallback) mask = MM_ACTIVE_NEEDS_GENERAL;
        }
        if (active) active[i] = (uint16_t)mask;
        if (iters) iters[i] = (uint8_t)(it > 255 ? 255 : it);
        if ((mask & MM_ACTIVE_FAILED) && n_fail) atomicAdd(n_fail, 1);
    }
};

// uncoalesced per-point accessor for the (rare) fix-up points, either layout
template <int LAYOUT> struct DirectAcc {
    const Ptrs<2, 3>& p;
    int64_t N, i;
    __device__ __forceinline__ DirectAcc(const Ptrs<2, 3>& p_, int64_t N_, int64_t i_) : p(p_), N(N_), i(i_) {}
    template <int A> __device__ __forceinline__ double in(int c) const {
        return LAYOUT == 0 ? p.in[A][(int64_t)c * N + i] : p.in[A][i * CrystalOp<0>::in_nc(A) + c];
    }
    template <int A> __device__ __forceinline__ void out(int c, double v) const {
        if (LAYOUT == 0) p.out[A][(int64_t)c * N + i] = v; else p.out[A][i * CrystalOp<0>::out_nc(A) + c] = v;
    }
    template <int A> __device__ __forceinline__ bool has_out() const { return p.out[A] != nullptr; }
};

template <int LAYOUT>
__global__ void __launch_bounds__(128) crystal_fixup_kernel(const CrystalK k, const Ptrs<2, 3> p, uint16_t* active, uint8_t* iters, int32_t* n_fail, const int64_t N) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= N) return;
    if (!(active[i] & MM_ACTIVE_NEEDS_GENERAL)) return;
    DirectAcc<LAYOUT> acc(p, N, i);
    int it = 0;
    const unsigned mask = crystal_point(k, acc, &it);
    active[i] = (uint16_t)mask;
    if (iters) iters[i] = (uint8_t)(it > 255 ? 255 : it);
    if ((mask & MM_ACTIVE_FAILED) && n_fail) atomicAdd(n_fail, 1);
}

#ifndef MM_CRYSTAL_BS
#define MM_CRYSTAL_BS 64
#endif
#ifndef MM_CRYSTAL_MINBLOCKS
#define MM_CRYSTAL_MINBLOCKS 4
#endif
template <> struct MinBlocks<CrystalOp<1>> { static constexpr int value = MM_CRYSTAL_MINBLOCKS; };
template <> struct MinBlocks<CrystalOp<2>> { static constexpr int value = MM_CRYSTAL_MINBLOCKS; };

template <int MODE>
int launch_crystal(const CrystalK& k, uint16_t* active, uint8_t* iters, int32_t* n_fail, const Ptrs<2, 3>& ptrs, int64_t N, int layout, cudaStream_t st) {
    CrystalOp<MODE> op;
    op.k = k; op.active = active; op.iters = iters; op.n_fail = n_fail;
    int rc = launch_points<CrystalOp<MODE>, (MODE == 0) ? 64 : MM_CRYSTAL_BS>(op, ptrs, N, layout, st, "mm_crystal");
    if (rc || MODE == 0 || N == 0) return rc;
    const int64_t blocks = (N + 127) / 128;
    if (layout == 0) crystal_fixup_kernel<0><<<(unsigned)blocks, 128, 0, st>>>(k, ptrs, active, iters, n_fail, N);
    else             crystal_fixup_kernel<1><<<(unsigned)blocks, 128, 0, st>>>(k, ptrs, active, iters, n_fail, N);
    return check_launch("mm_crystal (fix-up pass)");
}

}  // namespace mm

using namespace mm;

extern "C" int mm_crystal(const MMCrystal* p, const double* deps, double dt, const double* state_in, double* sig, double* tan,
                          double* state_out, uint16_t* active, uint8_t* iters, int32_t* n_fail, const MMNewtonOpts* opts,
                          int64_t N, int layout, void* stream) {
    if (!p || N < 0 || (N > 0 && (!deps || !state_in || !sig || !state_out))) return set_error(MM_ERR_ARG, "mm_crystal: bad argument");
    if (p->nslip != MM_NSLIP) return set_error(MM_ERR_ARG, "mm_crystal: nslip must be %d (got %d)", MM_NSLIP, p->nslip);
    CrystalK k;
    make_crystal_k(k, p->E, p->nu, p->tau_y, p->H_kin, p->alpha_inf, p->t_star, p->sigma_c, p->m, p->MS, p->H, dt,
                   opts ? opts->tol : 1e-8, opts ? opts->max_iter : 100);
    Ptrs<2, 3> ptrs; ptrs.in[0] = deps; ptrs.in[1] = state_in; ptrs.out[0] = sig; ptrs.out[1] = tan; ptrs.out[2] = state_out;
    cudaStream_t st = (cudaStream_t)stream;
    const char* force = getenv("MM_CRYSTAL_FORCE_GENERAL");      // debugging / A-B timing switch
    if (!k.h_struct || (force && force[0] == '1')) return launch_crystal<0>(k, active, iters, n_fail, ptrs, N, layout, st);
    // the structured path needs a per-point flag array for its fix-up pass; callers normally pass `active`,
    // otherwise a stream-ordered temporary is used (the only allocation this library ever makes on a compute call)
    uint16_t* flags = active;
    if (!flags && N > 0) {
        cudaError_t e = cudaMallocAsync((void**)&flags, (size_t)N * sizeof(uint16_t), st);
        if (e != cudaSuccess) return set_error(MM_ERR_CUDA, "mm_crystal: cudaMallocAsync: %s", cudaGetErrorString(e));
    }
    const int rc = (k.b_h != 0.0) ? launch_crystal<2>(k, flags, iters, n_fail, ptrs, N, layout, st)
                                  : launch_crystal<1>(k, flags, iters, n_fail, ptrs, N, layout, st);
    if (!active && flags) cudaFreeAsync(flags, st);
    return rc;
}
