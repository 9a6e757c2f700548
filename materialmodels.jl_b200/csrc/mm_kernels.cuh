// Generic point-wise kernel wrappers for sm_100a.
//
// Every model is an "Op": a POD functor carrying the material constants plus
//     static constexpr NIN / NOUT, in_nc(a) / out_nc(a)   (components per point of each array)
//     template<class Acc> __device__ void point(Acc& acc, int64_t i) const
// and is launched through one of two wrappers:
//   * soa_kernel — MM_LAYOUT_SOA, a[c*N + i]: thread <-> point, every load/store of a component is a
//     fully coalesced 256 B/warp streaming access (ld.global.cs / st.global.cs: data is touched once).
//   * aos_bulk_kernel — MM_LAYOUT_AOS, a[i*nc + c]: a tile of BS points of an array is one contiguous run in
//     global memory; it is staged through dense shared-memory images by 1-D TMA bulk copies (cp.async.bulk +
//     mbarrier in, bulk_group out) driven by one elected thread — no per-element copy instructions.
// Both paths run the same math (mm_math.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
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

template <class Op> struct MinBlocks { static constexpr int value = 1, block = 32; };   // per-Op occupancy target (register cap): `value` blocks of `block` threads

// SoA kernels that write the 36/81-component tangent keep 40+ output streams open per warp; with every warp slot of the SM filled
// (32-64 resident warps) the DRAM write stream thrashes — measured (profiles/r01_soa_residency_sweep.log): LinearElastic+tangent
// 0.944 -> 0.985 of the HBM peak and NeoHook 0.958 -> 0.980 when only 2 blocks of 256 threads are resident, while the read-dominated
// launches (σ only, energies, XuNeedleman) want full occupancy.  `blocks` > 0: cap residency at that many blocks per SM when out[1]
// (the tangent) is requested, by requesting unused dynamic shared memory.
template <class Op> struct SoaTangentCap { static constexpr int blocks = 0; };

// the AoS tile kernel may want a different register cap than the SoA kernel of the same op (it keeps fewer warps resident anyway:
// shared-memory images bound its residency); defaults to the SoA choice
template <class Op> struct MinBlocksAos { static constexpr int value = MinBlocks<Op>::value, block = MinBlocks<Op>::block; };

// Per-thread scratch in shared memory, laid out [slot][thread] (thread-minor => bank-conflict free).
// Ops declare how many doubles they need in Op::SCRATCH (0 for the streaming models).
struct Scratch {
    double* base;     // already offset by the thread index
    int stride;       // threads per block
    __device__ __forceinline__ double& operator()(int j) const { return base[j * stride]; }
};

// L2 prefetch hint of the SoA kernels.  A warp's loads of one component are two 128-byte lines; the kernels keep 8-12 warps per SM and have
// nothing else to do while a warp's 6-20 input lines travel (long-scoreboard is their top stall).  `pf` > 0: lanes 0 and 16 of every warp ask
// L2 for the input lines of the points `pf` further on — the block that becomes resident a fraction of a wave later — so that its loads find
// them on the chip.  No register, no scoreboard entry, no extra DRAM traffic (each line is still fetched once), no value changes.
// SoaPrefetch<Op>::dist(p) is the distance the launcher passes (0 = no hint); ops opt in where it was measured to pay (mm_ops.cu).
template <class Op> struct SoaPrefetch { static int dist(const Ptrs<Op::NIN, Op::NOUT>&) { return 0; } };
template <class Op, int A = 0> __device__ __forceinline__ void soa_prefetch_inputs(const Ptrs<Op::NIN, Op::NOUT>& p, int64_t N, int64_t j) {
    if constexpr (A < Op::NIN) {
#pragma unroll
        for (int c = 0; c < Op::in_nc(A); ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.in[A] + ((int64_t)c * N + j)));
        soa_prefetch_inputs<Op, A + 1>(p, N, j);
    }
}

template <class Op, int BS>
__global__ void __launch_bounds__(BS, (MinBlocks<Op>::value * MinBlocks<Op>::block + BS - 1) / BS) soa_kernel(const Op op, const Ptrs<Op::NIN, Op::NOUT> p, const int64_t N, const int pf) {
    extern __shared__ double sm_scratch[];
    const int64_t i = (int64_t)blockIdx.x * BS + threadIdx.x;
    if (i >= N) return;
    if (pf > 0 && (threadIdx.x & 15) == 0 && i + pf < N) soa_prefetch_inputs<Op>(p, N, i + pf);
    SoaAcc<Op::NIN, Op::NOUT> acc(p, N, i);
    Scratch scr{sm_scratch + threadIdx.x, BS};
    op.point(acc, i, scr);
}

// ---------------------------------------------------------------------------------------------
// AoS path (MM_LAYOUT_AOS): the copy engine moves whole tiles.  A tile of BS points of array A is one
// contiguous run of BS*nc doubles in global memory, so it is fetched with ONE 1-D bulk async copy (TMA,
// cp.async.bulk ... mbarrier::complete_tx) into a dense shared-memory image, and the outputs leave through bulk
// shared->global copies (bulk_group).  No per-element copy instructions, no address arithmetic; one elected thread
// drives the copies, the rest of the block waits on the mbarrier phase.  The next tile's loads are issued together
// with this tile's stores, so they overlap the store drain.  Bulk copies need 16-byte aligned bases and full tiles;
// the one ragged tile (or every tile, if a base is unaligned) uses element-wise copies inside the same kernel.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// dense tile images: inputs [A][pt][c], then outputs [A][pt][c] (no aliasing: strides differ)
// OutAlias<Op>::of(a) = index of the INPUT array whose image slot output a reuses, or -1.  A thread may overwrite its own point's
// input slots once it holds them in registers (no other thread reads them), so an output with the same component count as an input
// needs no image of its own: Plastic's σ(6)/state(14) go out through the ε(6)/state(14) slots and only the tangent is staged
// separately — 608 instead of 768 bytes of shared memory per point, one more resident block per SM.  Requires: the point function
// reads ALL inputs before its first output (program order keeps shared-memory loads ahead of the aliasing stores).
template <class Op> struct OutAlias { __host__ __device__ static constexpr int of(int) { return -1; } };
template <class Op> __host__ __device__ constexpr int dense_in_off(int a) { return a == 0 ? 0 : dense_in_off<Op>(a - 1) + Op::in_nc(a - 1); }
template <class Op> __host__ __device__ constexpr int dense_out_rel(int a) {     // offset inside the output image (aliased outputs take no room)
    return a == 0 ? 0 : dense_out_rel<Op>(a - 1) + (OutAlias<Op>::of(a - 1) < 0 ? Op::out_nc(a - 1) : 0);
}
template <class Op> __host__ __device__ constexpr int dense_out_rel_noalias(int a) { return a == 0 ? 0 : dense_out_rel_noalias<Op>(a - 1) + Op::out_nc(a - 1); }
template <class Op, int BS> struct DenseAcc {
    double* sm_in;            // input image  [A][pt][c] (also the staging area of aliased outputs)
    double* sm_out;           // output image [A][pt][c]
    const Ptrs<Op::NIN, Op::NOUT>& p;
    int t;
    __device__ __forceinline__ DenseAcc(double* in_, double* out_, const Ptrs<Op::NIN, Op::NOUT>& p_, int t_) : sm_in(in_), sm_out(out_), p(p_), t(t_) {}
    template <int A> __device__ __forceinline__ double in(int c) const { return sm_in[dense_in_off<Op>(A) * BS + t * Op::in_nc(A) + c]; }
    template <int A> __device__ __forceinline__ void out(int c, double v) const {
        constexpr int al = OutAlias<Op>::of(A);
        if constexpr (al >= 0) {
            static_assert(Op::out_nc(A) == Op::in_nc(al), "aliased output must have the input's component count");
            sm_in[dense_in_off<Op>(al) * BS + t * Op::out_nc(A) + c] = v;
        } else {
            sm_out[dense_out_rel<Op>(A) * BS + t * Op::out_nc(A) + c] = v;
        }
    }
    template <int A> __device__ __forceinline__ bool has_out() const { return p.out[A] != nullptr; }
};
// where output A is staged: in the input image (aliased) or in the output image
template <class Op, int BS, int A> __device__ __forceinline__ double* dense_out_ptr(double* sm_in, double* sm_out) {
    constexpr int al = OutAlias<Op>::of(A);
    if constexpr (al >= 0) return sm_in + dense_in_off<Op>(al) * BS;
    else return sm_out + dense_out_rel<Op>(A) * BS;
}
template <class Op, int BS, int A> struct BulkIO {
    static __device__ __forceinline__ uint32_t in_bytes() {
        uint32_t b = (uint32_t)(BS * Op::in_nc(A) * sizeof(double));
        if constexpr (A + 1 < Op::NIN) b += BulkIO<Op, BS, A + 1>::in_bytes();
        return b;
    }
    static __device__ __forceinline__ void load_all(double* sm_in, const Ptrs<Op::NIN, Op::NOUT>& p, int64_t i0, uint64_t* bar) {
        bulk_g2s(sm_in + dense_in_off<Op>(A) * BS, p.in[A] + i0 * Op::in_nc(A), (uint32_t)(BS * Op::in_nc(A) * sizeof(double)), bar);
        if constexpr (A + 1 < Op::NIN) BulkIO<Op, BS, A + 1>::load_all(sm_in, p, i0, bar);
    }
    static __device__ __forceinline__ void store_all(double* sm_in, double* sm_out, const Ptrs<Op::NIN, Op::NOUT>& p, int64_t i0) {
        if (p.out[A] != nullptr) bulk_s2g(p.out[A] + i0 * Op::out_nc(A), dense_out_ptr<Op, BS, A>(sm_in, sm_out), (uint32_t)(BS * Op::out_nc(A) * sizeof(double)));
        if constexpr (A + 1 < Op::NOUT) BulkIO<Op, BS, A + 1>::store_all(sm_in, sm_out, p, i0);
    }
};

template <class Op, int BS, int A> struct DenseTailIO {     // element-wise copies for the one ragged tile (npts < BS)
    static __device__ __forceinline__ void load_all(double* sm_in, const Ptrs<Op::NIN, Op::NOUT>& p, int64_t i0, int npts) {
        const double* g = p.in[A] + i0 * Op::in_nc(A);
        for (int e = threadIdx.x; e < npts * Op::in_nc(A); e += BS) sm_in[dense_in_off<Op>(A) * BS + e] = __ldcs(g + e);
        if constexpr (A + 1 < Op::NIN) DenseTailIO<Op, BS, A + 1>::load_all(sm_in, p, i0, npts);
    }
    static __device__ __forceinline__ void store_all(double* sm_in, double* sm_out, const Ptrs<Op::NIN, Op::NOUT>& p, int64_t i0, int npts) {
        if (p.out[A] != nullptr) {
            double* g = p.out[A] + i0 * Op::out_nc(A);
            const double* src = dense_out_ptr<Op, BS, A>(sm_in, sm_out);
            for (int e = threadIdx.x; e < npts * Op::out_nc(A); e += BS) __stcs(g + e, src[e]);
        }
        if constexpr (A + 1 < Op::NOUT) DenseTailIO<Op, BS, A + 1>::store_all(sm_in, sm_out, p, i0, npts);
    }
};

// `bulk` = 0 sends every tile through the element-wise copies (arrays whose base is not 16-byte aligned); the
// compute code is the same instantiation either way, so results do not depend on how a batch is tiled or chunked.
// NBUF input images: with 2, the loads of the NEXT tile are issued before this tile is computed (their DRAM latency hides behind
// the math instead of being exposed after every tile); with 1 they are issued together with this tile's stores.
#ifndef MM_AOS_NBUF
#define MM_AOS_NBUF 2
#endif
template <class Op, int BS> __host__ __device__ constexpr size_t aos_smem_bytes() {
    return 16 + (size_t)(MM_AOS_NBUF * dense_in_off<Op>(Op::NIN) + dense_out_rel<Op>(Op::NOUT) + Op::SCRATCH) * BS * sizeof(double);
}
template <class Op, int BS>
__global__ void __launch_bounds__(BS, (MinBlocksAos<Op>::value * MinBlocksAos<Op>::block + BS - 1) / BS)
aos_bulk_kernel(const Op op, const Ptrs<Op::NIN, Op::NOUT> p, const int64_t N, const int64_t ntiles, const int bulk) {
    constexpr int NB = MM_AOS_NBUF, IN_D = dense_in_off<Op>(Op::NIN) * BS, OUT_D = dense_out_rel<Op>(Op::NOUT) * BS;
    extern __shared__ __align__(128) unsigned char sm_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm_raw);               // NB barriers (<= 2) in the first 16 bytes
    double* sm_in = reinterpret_cast<double*>(sm_raw + 16);            // NB input images
    double* sm_out = sm_in + NB * IN_D;
    double* sm_scr = sm_out + OUT_D;
    if (threadIdx.x == 0) { for (int b = 0; b < NB; ++b) mbar_init(bar + b, 1); }
    __syncthreads();
    int64_t tile = blockIdx.x;
    if (tile >= ntiles) return;
    const int64_t nfull = bulk ? N / BS : 0;          // tiles [0, nfull) are full and go through the copy engine
    if (threadIdx.x == 0 && tile < nfull) {
        mbar_expect_tx(bar, BulkIO<Op, BS, 0>::in_bytes());
        BulkIO<Op, BS, 0>::load_all(sm_in, p, tile * BS, bar);
    }
    uint32_t k = 0;                                    // tiles processed by this CTA
    for (; tile < ntiles; tile += gridDim.x, ++k) {
        const bool full = tile < nfull;
        const int npts = (int)((N - tile * BS < BS) ? (N - tile * BS) : BS);
        const int buf = (NB == 2) ? (int)(k & 1u) : 0;
        double* in_img = sm_in + buf * IN_D;
        const int64_t nxt = tile + gridDim.x;
        if (NB == 2 && threadIdx.x == 0 && nxt < nfull) {             // image buf^1 was consumed one iteration ago (barrier at its end)
            mbar_expect_tx(bar + (buf ^ 1), BulkIO<Op, BS, 0>::in_bytes());
            BulkIO<Op, BS, 0>::load_all(sm_in + (buf ^ 1) * IN_D, p, nxt * BS, bar + (buf ^ 1));
        }
        if (full) {
            mbar_wait(bar + buf, (NB == 2) ? ((k >> 1) & 1u) : (k & 1u));   // this tile's inputs have landed
        } else {
            DenseTailIO<Op, BS, 0>::load_all(in_img, p, tile * BS, npts);
            __syncthreads();
        }
        if ((int)threadIdx.x < npts) {
            DenseAcc<Op, BS> acc(in_img, sm_out, p, (int)threadIdx.x);
            Scratch scr{sm_scr + threadIdx.x, BS};
            op.point(acc, tile * BS + threadIdx.x, scr);
        }
        fence_async_smem();                           // generic-proxy writes -> visible to the copy engine
        __syncthreads();                              // all outputs written, all inputs consumed
        if (full) {
            if (threadIdx.x == 0) {
                BulkIO<Op, BS, 0>::store_all(in_img, sm_out, p, tile * BS);
                bulk_commit();
                if (NB == 1 && nxt < nfull) {         // single image: next tile's loads overlap this tile's store drain
                    if (dense_out_rel<Op>(Op::NOUT) != dense_out_rel_noalias<Op>(Op::NOUT)) bulk_wait_read0();   // ... unless outputs leave from the input image
                    mbar_expect_tx(bar, BulkIO<Op, BS, 0>::in_bytes());
                    BulkIO<Op, BS, 0>::load_all(sm_in, p, nxt * BS, bar);
                }
                bulk_wait_read0();                    // output image may be overwritten again
            }
        } else {
            DenseTailIO<Op, BS, 0>::store_all(in_img, sm_out, p, tile * BS, npts);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// SoA arrays through the copy engine (compiled per Op through SoaTiled<Op>, selected per call): a tile of BS points of ONE component is a
// contiguous run of BS doubles, so a tile is NIN_comps + NOUT_comps bulk copies into / out of a component-major shared-memory image
// ([comp][point]: conflict-free for thread <-> point).  Built for the light ops with many output streams (XuNeedleman: 3 in, 12 out), whose
// direct kernel is bimodal in a sustained run — 1.85 ms per 1e8 points (1.00 of the HBM peak) or 2.2 ms.  Measured on five boxes
// (profiles/r02_xu_soa_tiled_sweeps.log): this kernel is bimodal as well (1.95-2.09 or 2.35-2.5 ms), i.e. no better in either
// mode, so the direct kernel stays the default and this one is an A/B option (mm_set_tuning("xu_soa_tiled", 1)).  Same double-buffered input
// scheme as aos_bulk_kernel.
#ifndef MM_SOA_TILE_ISSUERS
#define MM_SOA_TILE_ISSUERS 8      // warps whose lane 0 issues the copies of a tile (component runs dealt round-robin)
#endif
template <class Op> struct SoaTiled { static constexpr bool value = false; static constexpr int BS = 512; };
template <class Op, int BS> struct SoaTileAcc {
    double* sm_in;
    double* sm_out;
    const Ptrs<Op::NIN, Op::NOUT>& p;
    int t;
    __device__ __forceinline__ SoaTileAcc(double* in_, double* out_, const Ptrs<Op::NIN, Op::NOUT>& p_, int t_) : sm_in(in_), sm_out(out_), p(p_), t(t_) {}
    template <int A> __device__ __forceinline__ double in(int c) const { return sm_in[(dense_in_off<Op>(A) + c) * BS + t]; }
    template <int A> __device__ __forceinline__ void out(int c, double v) const { sm_out[(dense_out_rel_noalias<Op>(A) + c) * BS + t] = v; }
    template <int A> __device__ __forceinline__ bool has_out() const { return p.out[A] != nullptr; }
};
template <class Op, int BS, int A> struct SoaTileIO {
    static __device__ __forceinline__ void load_tail(double* sm_in, const Ptrs<Op::NIN, Op::NOUT>& p, int64_t N, int64_t i0, int npts) {
        for (int e = threadIdx.x; e < npts * Op::in_nc(A); e += BS) { const int c = e / npts, t = e - c * npts; sm_in[(dense_in_off<Op>(A) + c) * BS + t] = __ldcs(p.in[A] + (int64_t)c * N + i0 + t); }
        if constexpr (A + 1 < Op::NIN) SoaTileIO<Op, BS, A + 1>::load_tail(sm_in, p, N, i0, npts);
    }
    static __device__ __forceinline__ void store_tail(double* sm_out, const Ptrs<Op::NIN, Op::NOUT>& p, int64_t N, int64_t i0, int npts) {
        if (p.out[A] != nullptr)
            for (int e = threadIdx.x; e < npts * Op::out_nc(A); e += BS) { const int c = e / npts, t = e - c * npts; __stcs(p.out[A] + (int64_t)c * N + i0 + t, sm_out[(dense_out_rel_noalias<Op>(A) + c) * BS + t]); }
        if constexpr (A + 1 < Op::NOUT) SoaTileIO<Op, BS, A + 1>::store_tail(sm_out, p, N, i0, npts);
    }
};
// flat component index -> global run of that component (one elected lane per warp asks, so the compare chain is cheap)
template <class Op, int A = 0> __device__ __forceinline__ const double* soa_in_run(const Ptrs<Op::NIN, Op::NOUT>& p, int64_t N, int k) {
    if constexpr (A + 1 < Op::NIN) { if (k >= dense_in_off<Op>(A + 1)) return soa_in_run<Op, A + 1>(p, N, k); }
    return p.in[A] + (int64_t)(k - dense_in_off<Op>(A)) * N;
}
template <class Op, int A = 0> __device__ __forceinline__ double* soa_out_run(const Ptrs<Op::NIN, Op::NOUT>& p, int64_t N, int k) {
    if constexpr (A + 1 < Op::NOUT) { if (k >= dense_out_rel_noalias<Op>(A + 1)) return soa_out_run<Op, A + 1>(p, N, k); }
    return p.out[A] == nullptr ? nullptr : p.out[A] + (int64_t)(k - dense_out_rel_noalias<Op>(A)) * N;
}
template <class Op, int BS> __host__ __device__ constexpr size_t soa_tile_smem_bytes() {
    return 16 + (size_t)(2 * dense_in_off<Op>(Op::NIN) + dense_out_rel_noalias<Op>(Op::NOUT) + Op::SCRATCH) * BS * sizeof(double);
}
template <class Op, int BS>
__global__ void __launch_bounds__(BS, 1) soa_bulk_kernel(const Op op, const Ptrs<Op::NIN, Op::NOUT> p, const int64_t N, const int64_t ntiles, const int bulk) {
    constexpr int IN_C = dense_in_off<Op>(Op::NIN), OUT_C = dense_out_rel_noalias<Op>(Op::NOUT), IN_D = IN_C * BS, OUT_D = OUT_C * BS, NW = BS / 32;
    constexpr uint32_t run_bytes = (uint32_t)(BS * sizeof(double));
    extern __shared__ __align__(128) unsigned char sm_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm_raw);
    double* sm_in = reinterpret_cast<double*>(sm_raw + 16);
    double* sm_out = sm_in + 2 * IN_D;
    double* sm_scr = sm_out + OUT_D;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); }
    __syncthreads();
    int64_t tile = blockIdx.x;
    if (tile >= ntiles) return;
    const int64_t nfull = bulk ? N / BS : 0;
    constexpr uint32_t in_bytes = (uint32_t)(IN_D * sizeof(double));
    // the copies of a tile are spread over the warps (lane 0 of warp w moves component runs w, w + NW, ...): 15 copies issued from
    // one thread serialise on that thread's issue slot
    constexpr int ISSUERS = (MM_SOA_TILE_ISSUERS < NW) ? MM_SOA_TILE_ISSUERS : NW;
    const int warp = (int)(threadIdx.x >> 5);
    const bool leader = (threadIdx.x & 31u) == 0 && warp < ISSUERS;
    auto issue_loads = [&](double* img, int64_t i0, uint64_t* b) {
        if (threadIdx.x == 0) mbar_expect_tx(b, in_bytes);
        if (leader) for (int k = warp; k < IN_C; k += ISSUERS) bulk_g2s(img + k * BS, soa_in_run<Op>(p, N, k) + i0, run_bytes, b);
    };
    if (tile < nfull) issue_loads(sm_in, tile * BS, bar);
    uint32_t k = 0;
    for (; tile < ntiles; tile += gridDim.x, ++k) {
        const bool full = tile < nfull;
        const int npts = (int)((N - tile * BS < BS) ? (N - tile * BS) : BS);
        const int buf = (int)(k & 1u);
        double* in_img = sm_in + buf * IN_D;
        const int64_t nxt = tile + gridDim.x;
        if (nxt < nfull) issue_loads(sm_in + (buf ^ 1) * IN_D, nxt * BS, bar + (buf ^ 1));
        if (full) mbar_wait(bar + buf, (k >> 1) & 1u);
        else { SoaTileIO<Op, BS, 0>::load_tail(in_img, p, N, tile * BS, npts); __syncthreads(); }
        if ((int)threadIdx.x < npts) {
            SoaTileAcc<Op, BS> acc(in_img, sm_out, p, (int)threadIdx.x);
            Scratch scr{sm_scr + threadIdx.x, BS};
            op.point(acc, tile * BS + threadIdx.x, scr);
        }
        fence_async_smem();
        __syncthreads();
        if (full) {
            if (leader) {
                bool any = false;
                for (int c = warp; c < OUT_C; c += ISSUERS) {
                    double* g = soa_out_run<Op>(p, N, c);
                    if (g != nullptr) { bulk_s2g(g + tile * BS, sm_out + c * BS, run_bytes); any = true; }
                }
                if (any) { bulk_commit(); bulk_wait_read0(); }
            }
        } else {
            SoaTileIO<Op, BS, 0>::store_tail(sm_out, p, N, tile * BS, npts);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);   // mm_api.cu
int check_launch(const char* what);              // mm_api.cu
int sm_count();                                  // mm_api.cu
cudaError_t ws_alloc_async(void** ptr, size_t bytes, cudaStream_t st);   // mm_api.cu: stream-ordered workspace from the library's own pool

// opt-in to > 48 KB of dynamic shared memory: a per-function, PER-DEVICE attribute, so it is remembered per device
// (`done` is one static table per kernel instantiation) — a process that drives several GPUs stays correct
struct SmemOptIn { bool done[64] = {}; };
template <class K> int ensure_smem(SmemOptIn& st, K kernel, size_t bytes, const char* what) {
    if (bytes <= 48 * 1024) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 63;
    if (st.done[dev] && dev != 63) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return set_error(-2, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    st.done[dev] = true;
    return 0;
}

template <class Op, int BS_SOA, int BS_AOS = BS_SOA>
int launch_points(const Op& op, const Ptrs<Op::NIN, Op::NOUT>& p, int64_t N, int layout, cudaStream_t stream, const char* what, bool soa_tiled = false) {
    if (N == 0) return 0;
    if constexpr (SoaTiled<Op>::value) {
      if (layout == 0 && soa_tiled) {
        // copy-engine tiles for the ops that opt in; a run of one component must start 16-byte aligned for every component and tile:
        // even N and 16-byte aligned bases — otherwise (and for the ragged last tile) element-wise copies inside the same kernel
        constexpr int TBS = SoaTiled<Op>::BS;
        bool aligned = (N % 2 == 0);
        for (int a = 0; a < Op::NIN; ++a) aligned = aligned && ((reinterpret_cast<uintptr_t>(p.in[a]) & 15u) == 0);
        for (int a = 0; a < Op::NOUT; ++a) aligned = aligned && ((reinterpret_cast<uintptr_t>(p.out[a]) & 15u) == 0);
        constexpr size_t smem_b = soa_tile_smem_bytes<Op, TBS>();
        static SmemOptIn tile_optin;
        if (int rc = ensure_smem(tile_optin, soa_bulk_kernel<Op, TBS>, smem_b, what)) return rc;
        const int64_t ntiles = (N + TBS - 1) / TBS;
        static int resident = 0;                       // CTAs per SM that fit (shared-memory images bound it): one persistent wave
        if (resident == 0) {
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, soa_bulk_kernel<Op, TBS>, TBS, smem_b) != cudaSuccess || nb < 1) nb = 1;
            resident = nb;
        }
        int64_t blocks = (int64_t)sm_count() * resident;
        if (blocks > ntiles) blocks = ntiles;
        soa_bulk_kernel<Op, TBS><<<(unsigned)blocks, TBS, smem_b, stream>>>(op, p, N, ntiles, aligned ? 1 : 0);
        return check_launch(what);
      }
    }
    if (layout == 0) {
        const int64_t blocks = (N + BS_SOA - 1) / BS_SOA;
        if (blocks > 2147483647LL) return set_error(-1, "%s: N too large for one launch", what);
        constexpr size_t scratch = (size_t)Op::SCRATCH * BS_SOA * sizeof(double);
        // residency cap for the tangent-writing streaming ops (see SoaTangentCap): unused dynamic shared memory sized so that exactly
        // `cap` blocks fit the SM's 227 KB (+1 KB reserved per block)
        constexpr int cap = SoaTangentCap<Op>::blocks;
        constexpr size_t cap_bytes = cap > 0 ? (size_t)(227 * 1024 / cap - 1024 - 2048) : 0;
        const size_t smem = (cap > 0 && Op::NOUT > 1 && p.out[1] != nullptr && cap_bytes > scratch) ? cap_bytes : scratch;
        static SmemOptIn soa_optin;
        if (int rc = ensure_smem(soa_optin, soa_kernel<Op, BS_SOA>, cap_bytes > scratch ? cap_bytes : scratch, what)) return rc;
        soa_kernel<Op, BS_SOA><<<(unsigned)blocks, BS_SOA, smem, stream>>>(op, p, N, SoaPrefetch<Op>::dist(p));
    } else if (layout == 1) {
        // one kernel for every AoS batch: full tiles through the copy engine (TMA bulk copies) when all array bases are
        // 16-byte aligned, the ragged last tile (or everything, if unaligned) through element-wise copies
        bool aligned = true;
        for (int a = 0; a < Op::NIN; ++a) aligned = aligned && ((reinterpret_cast<uintptr_t>(p.in[a]) & 15u) == 0);
        for (int a = 0; a < Op::NOUT; ++a) aligned = aligned && ((reinterpret_cast<uintptr_t>(p.out[a]) & 15u) == 0);
        static int use_bulk = -1;
        if (use_bulk < 0) { const char* e = getenv("MM_AOS_BULK"); use_bulk = (e && e[0] == '0') ? 0 : 1; }
        constexpr size_t smem_b = aos_smem_bytes<Op, BS_AOS>();
        static SmemOptIn aos_optin;
        if (int rc = ensure_smem(aos_optin, aos_bulk_kernel<Op, BS_AOS>, smem_b, what)) return rc;
        const int64_t ntiles = (N + BS_AOS - 1) / BS_AOS;
        int64_t blocks = (int64_t)sm_count() * 16;          // grid-stride over tiles; 16 CTAs/SM covers any residency
        if (blocks > ntiles) blocks = ntiles;
        aos_bulk_kernel<Op, BS_AOS><<<(unsigned)blocks, BS_AOS, smem_b, stream>>>(op, p, N, ntiles, (aligned && use_bulk) ? 1 : 0);
    } else {
        return set_error(-1, "%s: unknown layout %d", what, layout);
    }
    return check_launch(what);
}

}  // namespace mm
