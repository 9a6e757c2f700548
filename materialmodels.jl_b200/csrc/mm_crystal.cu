// mm_crystal — device-pointer entry point of CrystalViscoPlastic{12} (FP64-compute-bound model).
#include <cuda_runtime.h>
#include "../../include/mmb200.h"
#include "mm_kernels.cuh"
#include "mm_crystal_math.cuh"

namespace mm {

struct CrystalOp {
    static constexpr int NIN = 2, NOUT = 3;
    __host__ __device__ static constexpr int in_nc(int a) { return a == 0 ? 6 : 42; }
    __host__ __device__ static constexpr int out_nc(int a) { return a == 0 ? 6 : (a == 1 ? 36 : 42); }
    CrystalK k;
    uint16_t* active;
    uint8_t* iters;
    int32_t* n_fail;
    template <class Acc> __device__ __forceinline__ void point(Acc& acc, int64_t i) const {
        int it = 0;
        const unsigned mask = crystal_point(k, acc, &it);
        if (active) active[i] = (uint16_t)mask;
        if (iters) iters[i] = (uint8_t)(it > 255 ? 255 : it);
        if ((mask & MM_ACTIVE_FAILED) && n_fail) atomicAdd(n_fail, 1);
    }
};

}  // namespace mm

using namespace mm;

extern "C" int mm_crystal(const MMCrystal* p, const double* deps, double dt, const double* state_in, double* sig, double* tan,
                          double* state_out, uint16_t* active, uint8_t* iters, int32_t* n_fail, const MMNewtonOpts* opts,
                          int64_t N, int layout, void* stream) {
    if (!p || N < 0 || (N > 0 && (!deps || !state_in || !sig || !state_out))) return set_error(MM_ERR_ARG, "mm_crystal: bad argument");
    if (p->nslip != MM_NSLIP) return set_error(MM_ERR_ARG, "mm_crystal: nslip must be %d (got %d)", MM_NSLIP, p->nslip);
    CrystalOp op;
    make_crystal_k(op.k, p->E, p->nu, p->tau_y, p->H_kin, p->alpha_inf, p->t_star, p->sigma_c, p->m, p->MS, p->H, dt,
                   opts ? opts->tol : 1e-8, opts ? opts->max_iter : 100);
    op.active = active; op.iters = iters; op.n_fail = n_fail;
    Ptrs<2, 3> ptrs; ptrs.in[0] = deps; ptrs.in[1] = state_in; ptrs.out[0] = sig; ptrs.out[1] = tan; ptrs.out[2] = state_out;
    return launch_points<CrystalOp, 64>(op, ptrs, N, layout, (cudaStream_t)stream, "mm_crystal");
}
