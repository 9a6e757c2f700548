// Generic point-wise kernel wrappers for sm_100a.
//
// Every model is an "Op": a POD functor carrying the material constants plus
//     static constexpr NIN / NOUT, in_nc(a) / out_nc(a)   (components per point of each array)
//     template<class Acc> __device__ void point(Acc& acc, int64_t i) const
// and is launched through one of two wrappers:
//   * soa_kernel — MM_LAYOUT_SOA, a[c*N + i]: thread <-> point, every load/store of a component is a
//     fully coalesced 256 B/warp streaming access (ld.global.cs / st.global.cs: data is touched once).
//   * aos_kernel — MM_LAYOUT_AOS, a[i*nc + c]: the block stages a tile of BS points per array through
//     shared memory with contiguous coalesced copies; per-thread accesses hit an odd-padded stride
//     (nc|1 doubles) so they are bank-conflict free.
// Both paths run the same math (mm_math.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "mm_math.cuh"

namespace mm {

template <int NIN, int NOUT> struct Ptrs {
    const double* in[NIN];
    double* out[NOUT];
};

// ---------------------------------------------------------------------------------------------
template <int NIN, int NOUT> struct SoaAcc {
    const Ptrs<NIN, NOUT>& p;
    int64_t N, i;
    __device__ __forceinline__ SoaAcc(const Ptrs<NIN, NOUT>& p_, int64_t N_, int64_t i_) : p(p_), N(N_), i(i_) {}
    template <int A> __device__ __forceinline__ double in(int c) const { return __ldcs(p.in[A] + ((int64_t)c * N + i)); }
    template <int A> __device__ __forceinline__ void out(int c, double v) const { __stcs(p.out[A] + ((int64_t)c * N + i), v); }
    template <int A> __device__ __forceinline__ bool has_out() const { return p.out[A] != nullptr; }
};

template <class Op> struct MinBlocks { static constexpr int value = 1; };   // per-Op occupancy target (register cap)

// Per-thread scratch in shared memory, laid out [slot][thread] (thread-minor => bank-conflict free).
// Ops declare how many doubles they need in Op::SCRATCH (0 for the streaming models).
struct Scratch {
    double* base;     // already offset by the thread index
    int stride;       // threads per block
    __device__ __forceinline__ double& operator()(int j) const { return base[j * stride]; }
};

template <class Op, int BS>
__global__ void __launch_bounds__(BS, MinBlocks<Op>::value) soa_kernel(const Op op, const Ptrs<Op::NIN, Op::NOUT> p, const int64_t N) {
    extern __shared__ double sm_scratch[];
    const int64_t i = (int64_t)blockIdx.x * BS + threadIdx.x;
    if (i >= N) return;
    SoaAcc<Op::NIN, Op::NOUT> acc(p, N, i);
    Scratch scr{sm_scratch + threadIdx.x, BS};
    op.point(acc, i, scr);
}

// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int pad_odd(int nc) { return nc | 1; }
template <class Op> __host__ __device__ constexpr int aos_in_off(int a) { return a == 0 ? 0 : aos_in_off<Op>(a - 1) + pad_odd(Op::in_nc(a - 1)); }
template <class Op> __host__ __device__ constexpr int aos_out_off(int a) { return a == 0 ? aos_in_off<Op>(Op::NIN) : aos_out_off<Op>(a - 1) + pad_odd(Op::out_nc(a - 1)); }
template <class Op> __host__ __device__ constexpr int aos_doubles_per_point() { return aos_out_off<Op>(Op::NOUT) + Op::SCRATCH; }

template <class Op, int BS> struct AosAcc {
    double* sm;
    const Ptrs<Op::NIN, Op::NOUT>& p;
    int t;
    __device__ __forceinline__ AosAcc(double* sm_, const Ptrs<Op::NIN, Op::NOUT>& p_, int t_) : sm(sm_), p(p_), t(t_) {}
    template <int A> __device__ __forceinline__ double in(int c) const { return sm[aos_in_off<Op>(A) * BS + t * pad_odd(Op::in_nc(A)) + c]; }
    template <int A> __device__ __forceinline__ void out(int c, double v) const { sm[aos_out_off<Op>(A) * BS + t * pad_odd(Op::out_nc(A)) + c] = v; }
    template <int A> __device__ __forceinline__ bool has_out() const { return p.out[A] != nullptr; }
};

template <int NC, int BS> __device__ __forceinline__ void tile_in(double* sm, const double* g, int npts) {
    constexpr int NCP = pad_odd(NC);
    const int n = npts * NC;
    for (int e = threadIdx.x; e < n; e += BS) {
        const int pt = e / NC, c = e - pt * NC;
        sm[pt * NCP + c] = __ldcs(g + e);
    }
}
template <int NC, int BS> __device__ __forceinline__ void tile_out(const double* sm, double* g, int npts) {
    constexpr int NCP = pad_odd(NC);
    const int n = npts * NC;
    for (int e = threadIdx.x; e < n; e += BS) {
        const int pt = e / NC, c = e - pt * NC;
        __stcs(g + e, sm[pt * NCP + c]);
    }
}
template <class Op, int BS, int A> struct TileIO {
    static __device__ __forceinline__ void load_all(double* sm, const Ptrs<Op::NIN, Op::NOUT>& p, int64_t i0, int npts) {
        tile_in<Op::in_nc(A), BS>(sm + aos_in_off<Op>(A) * BS, p.in[A] + i0 * Op::in_nc(A), npts);
        if constexpr (A + 1 < Op::NIN) TileIO<Op, BS, A + 1>::load_all(sm, p, i0, npts);
    }
    static __device__ __forceinline__ void store_all(const double* sm, const Ptrs<Op::NIN, Op::NOUT>& p, int64_t i0, int npts) {
        if (p.out[A] != nullptr) tile_out<Op::out_nc(A), BS>(sm + aos_out_off<Op>(A) * BS, p.out[A] + i0 * Op::out_nc(A), npts);
        if constexpr (A + 1 < Op::NOUT) TileIO<Op, BS, A + 1>::store_all(sm, p, i0, npts);
    }
};

template <class Op, int BS>
__global__ void __launch_bounds__(BS) aos_kernel(const Op op, const Ptrs<Op::NIN, Op::NOUT> p, const int64_t N, const int64_t ntiles) {
    extern __shared__ double sm[];
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t i0 = tile * BS;
        const int npts = (int)((N - i0 < BS) ? (N - i0) : BS);
        TileIO<Op, BS, 0>::load_all(sm, p, i0, npts);
        __syncthreads();
        if ((int)threadIdx.x < npts) {
            AosAcc<Op, BS> acc(sm, p, (int)threadIdx.x);
            Scratch scr{sm + aos_out_off<Op>(Op::NOUT) * BS + threadIdx.x, BS};
            op.point(acc, i0 + threadIdx.x, scr);
        }
        __syncthreads();
        TileIO<Op, BS, 0>::store_all(sm, p, i0, npts);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);   // mm_api.cu
int check_launch(const char* what);              // mm_api.cu
int sm_count();                                  // mm_api.cu

template <class Op, int BS>
int launch_points(const Op& op, const Ptrs<Op::NIN, Op::NOUT>& p, int64_t N, int layout, cudaStream_t stream, const char* what) {
    if (N == 0) return 0;
    if (layout == 0) {
        const int64_t blocks = (N + BS - 1) / BS;
        if (blocks > 2147483647LL) return set_error(-1, "%s: N too large for one launch", what);
        constexpr size_t scratch = (size_t)Op::SCRATCH * BS * sizeof(double);
        static bool soa_configured = false;
        if (!soa_configured && scratch > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(soa_kernel<Op, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scratch);
            if (e != cudaSuccess) return set_error(-2, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
            soa_configured = true;
        }
        soa_kernel<Op, BS><<<(unsigned)blocks, BS, scratch, stream>>>(op, p, N);
    } else if (layout == 1) {
        constexpr size_t smem = (size_t)aos_doubles_per_point<Op>() * BS * sizeof(double);
        static bool configured = false;
        if (!configured) {
            if (smem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(aos_kernel<Op, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return set_error(-2, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
            }
            configured = true;
        }
        const int64_t ntiles = (N + BS - 1) / BS;
        int64_t blocks = (int64_t)sm_count() * 8;
        if (blocks > ntiles) blocks = ntiles;
        aos_kernel<Op, BS><<<(unsigned)blocks, BS, smem, stream>>>(op, p, N, ntiles);
    } else {
        return set_error(-1, "%s: unknown layout %d", what, layout);
    }
    return check_launch(what);
}

}  // namespace mm
