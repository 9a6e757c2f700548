// mm_crystal — device-pointer entry point of CrystalViscoPlastic{12} (FP64-compute-bound model).
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/mmb200.h"
#include "mm_kernels.cuh"
#include "mm_crystal_math.cuh"

namespace mm {

// General path (any H; 12x12 pivoted LU in thread-local memory): one thread per point through the generic wrappers.
// Structured path (H = a I + b 11ᵀ, what the reference constructor builds): the two kernels below.  A point whose SPD
// factorisation or D-guard fails there is only MARKED (bit 14 of its mask); a third launch (crystal_fixup_kernel) re-runs
// exactly those points through the general pivoted solve, so the main kernels carry no fallback code (registers) and a rare
// fallback does not stall its warp.
#define MM_ACTIVE_NEEDS_GENERAL 0x4000u

struct CrystalGeneralOp {
    static constexpr int NIN = 2, NOUT = 3, SCRATCH = 144;     // the 12x12 Jacobian of each thread in its shared-memory scratch column
    __host__ __device__ static constexpr int in_nc(int a) { return a == 0 ? 6 : 42; }
    __host__ __device__ static constexpr int out_nc(int a) { return a == 0 ? 6 : (a == 1 ? 36 : 42); }
    CrystalK k;
    uint16_t* active;
    uint16_t* iters;
    int32_t* n_fail;
    template <class Acc> __device__ __forceinline__ void point(Acc& acc, int64_t i, const Scratch& scr) const {
        int it = 0;
        const unsigned mask = crystal_point(k, acc, &it, scr.base, scr.stride);
        if (active) active[i] = (uint16_t)mask;
        if (iters) iters[i] = (uint16_t)(it > 65535 ? 65535 : it);
        if ((mask & MM_ACTIVE_FAILED) && n_fail) atomicAdd(n_fail, 1);
    }
};

// uncoalesced per-point accessor for the (rare) fix-up points, either layout
// `ii` = where the point's inputs live, `i` = where its outputs go: the same index, except for the index-list batches of the dimension
// wrappers (inputs stay where the point is, outputs are packed by list position: dense, coalesced stores whatever the list looks like)
template <int LAYOUT> struct DirectAcc {
    const Ptrs<2, 3>& p;
    int64_t N, i, ii;
    __device__ __forceinline__ DirectAcc(const Ptrs<2, 3>& p_, int64_t N_, int64_t i_) : p(p_), N(N_), i(i_), ii(i_) {}
    __device__ __forceinline__ DirectAcc(const Ptrs<2, 3>& p_, int64_t N_, int64_t i_, int64_t ii_) : p(p_), N(N_), i(i_), ii(ii_) {}
    template <int A> __device__ __forceinline__ double in(int c) const {
        return LAYOUT == 0 ? p.in[A][(int64_t)c * N + ii] : p.in[A][ii * CrystalGeneralOp::in_nc(A) + c];
    }
    template <int A> __device__ __forceinline__ void out(int c, double v) const {
        if (LAYOUT == 0) p.out[A][(int64_t)c * N + i] = v; else p.out[A][i * CrystalGeneralOp::out_nc(A) + c] = v;
    }
    template <int A> __device__ __forceinline__ bool has_out() const { return p.out[A] != nullptr; }
};

// Fix-up pass: the flagged points are few and scattered (none for FCC; ~2 % on the reference's BCC table), so running them where they
// lie would leave 31 of 32 lanes idle through a 12x12 pivoted-LU Newton solve (measured: BCC 415 ms per 1e7 points against 25 ms for
// FCC).  Each block scans a contiguous range of the flag array, compacts the flagged indices into shared memory, and then works the
// list off with full lanes.
#define MM_FIXUP_BS 32
#define MM_FIXUP_RANGE 4096
template <int LAYOUT>
__global__ void __launch_bounds__(MM_FIXUP_BS) crystal_fixup_kernel(const CrystalK k, const Ptrs<2, 3> p, uint16_t* active, uint16_t* iters, int32_t* n_fail, const int64_t N,
                                                                    const int64_t nranges, const int* cnt, const int* idx, const int* n_general) {
    if (n_general && *n_general == 0) return;                             // nothing was flagged (the usual case)
    __shared__ uint16_t list[MM_FIXUP_RANGE];
    __shared__ int count, next;
    __shared__ double Jsm[144 * MM_FIXUP_BS];                             // every thread's 12x12 Jacobian: column t of a [144][BS] image
    for (int64_t r = blockIdx.x; r < nranges; r += gridDim.x) {           // persistent blocks: few resident threads keep the per-thread
        __syncthreads();                                                  // solver state (2.4 KB of local memory each) inside L1
        if (threadIdx.x == 0) { count = 0; next = 0; }
        __syncthreads();
        const int64_t base = r * MM_FIXUP_RANGE;
        const int64_t Nc = cnt ? (int64_t)*cnt : N;                       // points in use (cnt: device-side count; N stays the stride)
        const int n = (int)((Nc - base < MM_FIXUP_RANGE) ? (Nc - base) : MM_FIXUP_RANGE);
        for (int j = threadIdx.x; j < n; j += MM_FIXUP_BS)
            if (active[base + j] & MM_ACTIVE_NEEDS_GENERAL) list[atomicAdd(&count, 1)] = (uint16_t)j;
        __syncthreads();
        const int m = count;
        if (m == 0) continue;
        // lanes take list entries one at a time and advance ONE Newton pass per trip, so a lane whose solve ends (converged after a few
        // passes) moves on to the next point while its neighbours keep iterating: a non-converging point no longer holds 31 others for
        // max_iter passes
        CrystalGen S;
        bool has = false;
        int64_t i = 0;
        for (;;) {
            if (!has) {
                const int e = atomicAdd(&next, 1);
                if (e < m) {
                    i = base + list[e];
                    DirectAcc<LAYOUT> acc(p, N, i, idx ? (int64_t)idx[i] : i);
                    crystal_gen_init(k, acc, S, Jsm + threadIdx.x, MM_FIXUP_BS);
                    has = true;
                }
            }
            if (!__any_sync(0xffffffffu, has)) break;
            if (has && crystal_gen_pass(k, S)) {
                DirectAcc<LAYOUT> acc(p, N, i, idx ? (int64_t)idx[i] : i);
                const unsigned mask = crystal_gen_finalize(k, acc, S);
                active[i] = (uint16_t)mask;
                if (iters) iters[i] = (uint16_t)(S.iters > 65535 ? 65535 : S.iters);
                if ((mask & MM_ACTIVE_FAILED) && n_fail) atomicAdd(n_fail, 1);
                has = false;
            }
        }
    }
}

int* wrap_tuning_slot(const char* name);      // mm_wrap.cu
int* ops_tuning_slot(const char* name);       // mm_ops.cu
int* host_tuning_slot(const char* name);      // mm_host.cu
// tuning / A-B switches: initialised from the environment ONCE per process (tools/README.md), changed afterwards only through
// mm_set_tuning() — never a getenv on the per-call path
struct CrystalKnobs { int thresh, chunk, fixup_per_sm, force_general, variant, finish_pf; };
static CrystalKnobs& crystal_knobs() {
    static CrystalKnobs k = [] {
        CrystalKnobs v{0, 0, 0, 0, -1, 4096};       // finish_pf: 17.00-17.02 -> 16.82-16.85 ms per 1e7 points at 2 048-4 096 points ahead, no gain from 16 384 on (profiles/r02_crystal_finish_prefetch*.log)
        if (const char* e = getenv("MM_CRYSTAL_THRESH")) v.thresh = atoi(e);
        if (const char* e = getenv("MM_CRYSTAL_CHUNK")) v.chunk = atoi(e);
        if (const char* e = getenv("MM_FIXUP_PER_SM")) v.fixup_per_sm = atoi(e);
        if (const char* e = getenv("MM_CRYSTAL_FORCE_GENERAL")) v.force_general = (e[0] == '1');
        if (const char* e = getenv("MM_CRYSTAL_VARIANT")) v.variant = atoi(e);
        if (const char* e = getenv("MM_CRYSTAL_FINISH_PREFETCH")) v.finish_pf = atoi(e);
        return v;
    }();
    return k;
}

#ifndef MM_CRYSTAL_BS
#define MM_CRYSTAL_BS 64
#endif
#ifndef MM_CRYSTAL_MINBLOCKS
#define MM_CRYSTAL_MINBLOCKS 6
#endif
#ifndef MM_CRYSTAL_THRESH
#define MM_CRYSTAL_THRESH 4
#endif

// ---------------------------------------------------------------------------------------------------------
// Two-kernel structured path.
// (1) crystal_newton_kernel — lane-persistent Newton.  Iteration counts differ wildly between neighbouring points
//     (0 for an elastic point, 5-30 for a plastic one), so with one point per thread a warp idles most of its lanes
//     while the slowest solve finishes.  Here each WARP owns a contiguous chunk of points; a lane whose solve is
//     finished stores its converged μ (into the μ slots of the state output), its iteration count and a status word,
//     and is refilled with the next point of the chunk — in batches of >= `thresh` lanes so that the (divergent)
//     store + refill code runs with many lanes; warp-level voting (__ballot_sync) decides every phase.  All
//     per-point vectors live in the lane's shared-memory scratch column.
// (2) crystal_finish_kernel — one thread per point, no divergence between neighbours: re-evaluates at the stored μ
//     (one extra pass out of ~15) and writes σ, the tangent (six 6x6 solves) and the state.  In the one-kernel version
//     this phase ran with 5.8 of 32 lanes active and took 16 % of the time (profiles/r01_crystal_v5_ncu_full.txt).
// ---------------------------------------------------------------------------------------------------------
// The Newton kernel hands its converged μ to the finish kernel through the μ slots of the state OUTPUT — unless the caller aliased
// state_in and state_out (allowed by include/mmb200.h): then those slots still hold the PREVIOUS μ, which the fix-up pass needs as the
// reference's initial guess x0 = state.μ (CrystalViscoPlastic.jl:73) for the points the finish kernel punts to it, so the hand-over goes
// through a stream-ordered side buffer xbuf[12][N] instead and nothing of the old state is overwritten before the fix-up decision.
template <int LAYOUT> struct DirectAccRW : DirectAcc<LAYOUT> {
    const double* xbuf;
    __device__ __forceinline__ DirectAccRW(const Ptrs<2, 3>& p_, int64_t N_, int64_t i_, int64_t ii_, const double* xbuf_) : DirectAcc<LAYOUT>(p_, N_, i_, ii_), xbuf(xbuf_) {}
    template <int A> __device__ __forceinline__ double rd_out(int c) const {
        if (xbuf) return xbuf[(int64_t)(c - 30) * this->N + this->i];
        return LAYOUT == 0 ? this->p.out[A][(int64_t)c * this->N + this->i] : this->p.out[A][this->i * CrystalGeneralOp::out_nc(A) + c];
    }
};
template <int LAYOUT, int XSLOT = XS, class Scr> __device__ __forceinline__ void crystal_hand_over(const Ptrs<2, 3>& p, double* xbuf, int64_t N, int64_t my, const Scr& scr) {
    if (xbuf) {
#pragma unroll
        for (int i = 0; i < 12; ++i) xbuf[(int64_t)i * N + my] = scr(XSLOT + i);
    } else {
        DirectAcc<LAYOUT> acc(p, N, my);
#pragma unroll
        for (int i = 0; i < 12; ++i) acc.template out<2>(30 + i, scr(XSLOT + i));
    }
}

// chunk of points per warp: large enough to refill many times, small enough for >= ~8 waves of warps on the chip (rw8 = 8 x resident warps)
__host__ __device__ inline int crystal_chunk_for(int64_t n, int rw8) {
    int64_t c = n / (rw8 > 0 ? rw8 : 1);
    if (c > 1024) c = 1024;
    if (c < 64) c = 64;
    return (int)((c + 31) / 32 * 32);
}

// (0) crystal_presort_kernel — scheduling only, no effect on any result.  The sparse pass costs a warp its LONGEST per-lane candidate list,
//     and neighbouring points have unrelated lists (configs[3]: 3.0 candidates per lane-pass on average, 7 for the longest of 32), so 15 of 32
//     lanes are busy.  The candidate count of a point's FIRST pass predicts its mean count over all passes (correlation 0.91 on configs[3],
//     tools/crystal_sched_sim.py), so each chunk (the unit one Newton warp owns) is re-ordered by that count, heaviest first: the lanes of a warp
//     then walk lists of similar length (model: 0.39 -> 0.63 of the issued lane-slots useful, 26 % fewer warp-instructions), and the points that
//     end after one evaluation come last, where they shorten the drain of the chunk.  One block per chunk; stable counting sort on 13 keys
//     (ballot ranks inside 32-point segments, serial prefix over the <= 32 segments); order[chunk_base + position] = offset of the point in its chunk.
#define MM_PRESORT_BS 256
#define MM_ASET_MAX 6      // longest candidate list the active-set pass solves (crystal_iterate_aset)
template <bool HASQ, int LAYOUT>
__global__ void __launch_bounds__(MM_PRESORT_BS) crystal_presort_kernel(const CrystalK k, const Ptrs<2, 3> p, uint16_t* order, const int64_t N, const int chunk, uint16_t* hc, int* heavy_idx, int* heavy_cnt) {
    __shared__ uint16_t segcnt[32][13];         // [segment of 32 points][key]  -> exclusive prefix over the segments
    __shared__ uint16_t binstart[13];
    __shared__ int hcount, hoff;                // points of this chunk with more than MM_ASET_MAX candidates, and where their indices go in the global list
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * chunk;
    const int n = (int)((N - base < chunk) ? (N - base) : chunk);
    for (int e = threadIdx.x; e < 32 * 13; e += MM_PRESORT_BS) (&segcnt[0][0])[e] = 0;
    __syncthreads();
    int key[4], rank[4];                       // chunk <= 1024 = 4 rounds of 256 threads
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int j = r * MM_PRESORT_BS + threadIdx.x;
        key[r] = -1; rank[r] = 0;
        if (j - lane >= n) continue;           // whole segment beyond the chunk (warp-uniform)
        int kk = 13;                           // lanes beyond the end: a key of their own
        if (j < n) {
            DirectAcc<LAYOUT> acc(p, N, base + j);
            const double G2 = 2.0 * k.G;
            double e6[6], sg[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) e6[c] = acc.template in<0>(c);
            const double ltr = k.lam * tr6(e6);
#pragma unroll
            for (int c = 0; c < 6; ++c) sg[c] = acc.template in<1>(c) + (is_diag_slot(c) ? (G2 * e6[c] + ltr) : (G2 * e6[c]));
            double st[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) st[c] = sg[c];
            unsigned nz = 0;
            double sumx = 0.0;
            for (int i = 0; i < 12; ++i) {                         // σ(x0) = σ_trial − Σ x0_i s_i Eᵉ:MS_i, s_i from the trial state
                const double x0 = acc.template in<1>(30 + i);
                if (x0 != 0.0) {
                    nz |= 1u << i;
                    sumx += x0;
                    double tau = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) tau += st[c] * k.Pw[i][c];
                    const double d = tau - acc.template in<1>(18 + i);
                    const double f = x0 * ((d > 0.0) ? 1.0 : ((d < 0.0) ? -1.0 : 0.0));
#pragma unroll
                    for (int c = 0; c < 6; ++c) sg[c] -= f * k.EM[i][c];
                }
            }
            unsigned cand = nz;
            for (int i = 0; i < 12; ++i) {
                double tau = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) tau += sg[c] * k.Pw[i][c];
                double kap = acc.template in<1>(6 + i);
                if (HASQ) kap += k.b_h * sumx;
                const double Phi = fabs(tau - acc.template in<1>(18 + i)) - k.tau_y - kap;
                cand |= (!(Phi <= 0.0)) ? (1u << i) : 0u;
            }
            kk = 12 - __popc(cand);                                // heaviest first
        }
        const unsigned same = __match_any_sync(0xffffffffu, kk);
        key[r] = kk;
        rank[r] = __popc(same & ((1u << lane) - 1u));
        if (kk < 13 && rank[r] == 0) segcnt[r * (MM_PRESORT_BS / 32) + warp][kk] = (uint16_t)__popc(same);
    }
    __syncthreads();
    if (threadIdx.x < 13) {                    // exclusive prefix over the segments, per key
        int run = 0;
        for (int sgm = 0; sgm < 32; ++sgm) { const int c = segcnt[sgm][threadIdx.x]; segcnt[sgm][threadIdx.x] = (uint16_t)run; run += c; }
        binstart[threadIdx.x] = (uint16_t)run;                      // total of this key, for now
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int q = 0; q < 13; ++q) { const int c = binstart[q]; binstart[q] = (uint16_t)run; run += c; }
        // the active-set launch starts each chunk behind its heavy points (more than MM_ASET_MAX candidates in the first pass: keys 0 .. 11 − MAX);
        // their indices go to the global list the sparse launch works off
        hcount = hc ? binstart[12 - MM_ASET_MAX] : 0;
        if (hc) { hc[blockIdx.x] = (uint16_t)hcount; hoff = hcount ? atomicAdd(heavy_cnt, hcount) : 0; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int j = r * MM_PRESORT_BS + threadIdx.x;
        if (key[r] >= 0 && key[r] < 13) {
            const int pos = binstart[key[r]] + segcnt[r * (MM_PRESORT_BS / 32) + warp][key[r]] + rank[r];
            order[base + pos] = (uint16_t)j;
            if (pos < hcount) heavy_idx[hoff + pos] = (int)(base + j);
        }
    }
}

// scheduling data of one Newton launch (all optional): the chunk-local order, the per-chunk count of heavy points the launch skips, the global list
// that collects points handed over to the sparse launch, and `scatter` = the launch works off such a list (idx / cnt) and writes each point's
// results at the listed index (a list of the wrappers' batch POSITIONS then: `in_idx` = the batch's own list says where each position's inputs are)
struct CrystalSched { const uint16_t* order; const uint16_t* hc; int* heavy_idx; int* heavy_cnt; int scatter; int* n_general; const int* in_idx; };

// SPARSE (integer m only): the per-lane active-list pass crystal_iterate_sparse — the slip-system tables are copied into shared memory
// ([component][system], behind the scratch columns) because every lane walks its own list of systems
#ifndef MM_CRYSTAL_ASET_MINBLOCKS
#define MM_CRYSTAL_ASET_MINBLOCKS 8   // 16 resident warps (128 registers, 48 scratch slots per lane): 12.40 ms per 1e7 points against 13.14 (6 blocks) / 12.74 (7), profiles/r02_crystal_aset_occupancy.log
#endif
template <bool HASQ, bool INTPOW, int LAYOUT, int BS, int PASS>      // PASS: 0 dense, 1 sparse, 2 active-set (q = 0; its own, smaller scratch layout)
__global__ void __launch_bounds__(BS, PASS == 2 ? (HASQ ? MM_CRYSTAL_ASET_MINBLOCKS - 1 : MM_CRYSTAL_ASET_MINBLOCKS) : MM_CRYSTAL_MINBLOCKS)      // (q != 0: 54 scratch slots, 7 blocks)
crystal_newton_kernel(const CrystalK k, const Ptrs<2, 3> p, uint16_t* active, uint16_t* iters, double* xbuf, const int64_t N, int chunk, const int thresh, const int* cnt, const int* idx, const int rw8,
                      const CrystalSched sched) {
    extern __shared__ double sm_scratch[];
    const Scratch scr{sm_scratch + threadIdx.x, BS};
    double* tab = sm_scratch + (PASS == 2 ? (HASQ ? MM_CRYSTAL_SCRATCH_ASET_Q : MM_CRYSTAL_SCRATCH_ASET) : MM_CRYSTAL_SCRATCH) * BS;
    double sigt[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};          // active-set pass: σ_trial of the lane's point
    if (PASS != 0) {
        for (int e = threadIdx.x; e < MM_CRYSTAL_TAB; e += BS) crystal_fill_tab(k, tab, e);
        __syncthreads();
    }
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t warp_id = (int64_t)blockIdx.x * (BS / 32) + (threadIdx.x >> 5);
    const int64_t Nc = cnt ? (int64_t)*cnt : N;               // points in use (N stays the SoA stride)
    const int64_t nwarps = (int64_t)gridDim.x * (BS / 32);
    if (cnt) chunk = crystal_chunk_for(Nc, rw8);              // batch: the chunk follows the device-side count (the grid covers N / 64 warps)
    if (cnt && sched.scatter) {                               // one resident wave works the whole list: one chunk per warp, so that its lanes drain once
        const int64_t share = ((Nc + nwarps - 1) / nwarps + 31) / 32 * 32;
        if (share > chunk) chunk = (int)(share < (1 << 20) ? share : (1 << 20));
    }
    CrystalLane L;
    for (int64_t cid = warp_id; cid * chunk < Nc; cid += nwarps) {   // one trip, except for a scatter launch (one resident wave walks all chunks of its list)
    const int64_t chunk0 = cid * chunk;
    int64_t next = chunk0 + (sched.hc ? (int64_t)sched.hc[cid] : 0);   // the chunk's heavy points (sorted first) belong to the sparse launch
    const int64_t end = (chunk0 + chunk < Nc) ? (chunk0 + chunk) : Nc;
    bool has_work = false, finished = false;
    int64_t my = 0;
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !has_work);
        if (idle != 0u && next < end) {                       // refill idle lanes with the next points of the chunk
            if (!has_work) {
                const int64_t cand = next + __popc(idle & ((1u << lane) - 1u));
                if (cand < end) {
                    // output slot; the inputs of a wrapper's list entry stay where the point is (order: the chunk's points, heaviest first)
                    my = sched.scatter ? (int64_t)idx[cand] : (sched.order ? (chunk0 + sched.order[cand]) : cand);
                    DirectAcc<LAYOUT> acc(p, N, my, sched.scatter ? (sched.in_idx ? (int64_t)sched.in_idx[my] : my) : (idx ? (int64_t)idx[cand] : my));
                    if (PASS == 2) crystal_init_t<XSA, -1>(k, acc, scr, L, sigt);
                    else           crystal_init(k, acc, scr, L);
                    has_work = true; finished = false;
                }
            }
            next += __popc(idle);
        }
        const unsigned working = __ballot_sync(FULL, has_work);
        if (working == 0u) break;
        if (PASS == 2) {
            const unsigned lanes = __ballot_sync(FULL, has_work && !finished);
            if (has_work && !finished) finished = crystal_iterate_aset<HASQ, INTPOW>(k, tab, scr, L, sigt, lanes);
        } else if (has_work && !finished) {
            if (PASS == 1) finished = crystal_iterate_sparse<HASQ, false, INTPOW>(k, tab, scr, L);
            else           finished = crystal_iterate<HASQ, INTPOW, false>(k, scr, L);
        }
        const unsigned fin = __ballot_sync(FULL, has_work && finished);
        const unsigned iterating = working & ~fin;
        if (fin != 0u && (__popc(fin) >= thresh || iterating == 0u || next >= end)) {
            if (has_work && finished) {
                if (PASS == 2 && L.big) {
                    sched.heavy_idx[atomicAdd(sched.heavy_cnt, 1)] = (int)my;      // redone from its first pass by the sparse launch; nothing written
                } else {
                    if (!L.bad) crystal_hand_over<LAYOUT, (PASS == 2 ? XSA : XS)>(p, xbuf, N, my, scr);
                    else if (sched.n_general) atomicAdd(sched.n_general, 1);
                    active[my] = (uint16_t)(L.bad ? MM_ACTIVE_NEEDS_GENERAL : (L.converged ? 0u : MM_ACTIVE_FAILED));
                    if (iters) iters[my] = (uint16_t)(L.bad ? 0 : (L.iters > 65535 ? 65535 : L.iters));
                }
                has_work = false; finished = false;
            }
        }
    }
    }
}

// (1') the same lane-persistent scheme around the PHASED pass (crystal_iterate_phased: integer m, q = 0 — the reference's test and
// BASELINE configuration).  Its own kernel because its parameter block is its own: every table entry is read as a compile-time
// constant-bank operand, so the block holds exactly the tables the pass touches (Pw, Eᵉ:MS, T, (Eᵉ)⁻¹: 3.4 KB of the 4 KB bank window).
#ifndef MM_CRYSTAL_PHASED_MINBLOCKS
#define MM_CRYSTAL_PHASED_MINBLOCKS 6
#endif
template <int LAYOUT, int BS>
__global__ void __launch_bounds__(BS, MM_CRYSTAL_PHASED_MINBLOCKS)
crystal_newton_phased_kernel(const CrystalNK k, const Ptrs<2, 3> p, uint16_t* active, uint16_t* iters, double* xbuf, const int64_t N, int chunk, const int thresh, const int* cnt, const int* idx, const int rw8, int* n_general) {
    extern __shared__ double sm_scratch[];
    const Scratch scr{sm_scratch + threadIdx.x, BS};
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t warp_id = (int64_t)blockIdx.x * (BS / 32) + (threadIdx.x >> 5);
    const int64_t Nc = cnt ? (int64_t)*cnt : N;
    if (cnt) chunk = crystal_chunk_for(Nc, rw8);
    int64_t next = warp_id * chunk;
    const int64_t end = (next + chunk < Nc) ? (next + chunk) : Nc;
    CrystalLane L;
    double sigt[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    bool has_work = false, finished = false;
    int64_t my = 0;
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !has_work);
        if (idle != 0u && next < end) {                       // refill idle lanes with the next points of the chunk
            if (!has_work) {
                const int64_t cand = next + __popc(idle & ((1u << lane) - 1u));
                if (cand < end) {
                    my = cand;                                   // output slot; the inputs of a list entry stay where the point is
                    DirectAcc<LAYOUT> acc(p, N, my, idx ? (int64_t)idx[cand] : cand);
                    crystal_init(k, acc, scr, L);
#pragma unroll
                    for (int c = 0; c < 6; ++c) sigt[c] = scr(SIGT + c);
                    has_work = true; finished = false;
                }
            }
            next += __popc(idle);
        }
        const unsigned working = __ballot_sync(FULL, has_work);
        if (working == 0u) break;
        if (has_work && !finished) finished = crystal_iterate_phased<false>(k, scr, L, sigt);
        const unsigned fin = __ballot_sync(FULL, has_work && finished);
        const unsigned iterating = working & ~fin;
        if (fin != 0u && (__popc(fin) >= thresh || iterating == 0u || next >= end)) {
            if (has_work && finished) {
                if (!L.bad) crystal_hand_over<LAYOUT>(p, xbuf, N, my, scr);
                else if (n_general) atomicAdd(n_general, 1);
                active[my] = (uint16_t)(L.bad ? MM_ACTIVE_NEEDS_GENERAL : (L.converged ? 0u : MM_ACTIVE_FAILED));
                if (iters) iters[my] = (uint16_t)(L.bad ? 0 : (L.iters > 65535 ? 65535 : L.iters));
                has_work = false; finished = false;
            }
        }
    }
}

#ifndef MM_CRYSTAL_FINISH_BS
#define MM_CRYSTAL_FINISH_BS 64
#endif
#ifndef MM_CRYSTAL_FINISH_MINBLOCKS
#define MM_CRYSTAL_FINISH_MINBLOCKS 4
#endif
template <bool HASQ, bool INTPOW, int LAYOUT, int BS>
__global__ void __launch_bounds__(BS, MM_CRYSTAL_FINISH_MINBLOCKS)
crystal_finish_kernel(const CrystalK k, const Ptrs<2, 3> p, uint16_t* active, int32_t* n_fail, const double* xbuf, const int64_t N, const int* cnt, const int* idx, int* n_general, const int pf) {
    extern __shared__ double sm_scratch[];
    const Scratch scr{sm_scratch + threadIdx.x, BS};
    const int64_t i = (int64_t)blockIdx.x * BS + threadIdx.x;
    if (i >= (cnt ? (int64_t)*cnt : N)) return;
    if (LAYOUT == 0 && pf > 0 && !idx && (threadIdx.x & 15) == 0 && i + pf < N) {      // L2 hint for the rows of the points `pf` further on ("crystal_finish_prefetch"; see soa_kernel)
        const int64_t j = i + pf;
#pragma unroll
        for (int c = 0; c < 6; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.in[0] + (int64_t)c * N + j));
#pragma unroll
        for (int c = 0; c < 42; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.in[1] + (int64_t)c * N + j));
#pragma unroll
        for (int c = 0; c < 12; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(xbuf ? xbuf + (int64_t)c * N + j : p.out[2] + (int64_t)(30 + c) * N + j));
    }
    const unsigned status = active[i];
    if (status & MM_ACTIVE_NEEDS_GENERAL) return;                     // left to the fix-up pass
    DirectAccRW<LAYOUT> acc(p, N, i, idx ? (int64_t)idx[i] : i, xbuf);
    bool fallback = false;
    unsigned mask = crystal_finish_from_x<HASQ, INTPOW>(k, acc, scr, (status & MM_ACTIVE_FAILED) == 0, &fallback);
    if (fallback) { mask = MM_ACTIVE_NEEDS_GENERAL; if (n_general) atomicAdd(n_general, 1); }
    active[i] = (uint16_t)mask;
    if ((mask & MM_ACTIVE_FAILED) && n_fail) atomicAdd(n_fail, 1);
}

template <bool HASQ, bool INTPOW, int LAYOUT>
int launch_crystal_two_kernel(const CrystalK& k, uint16_t* active, uint16_t* iters, int32_t* n_fail, const Ptrs<2, 3>& ptrs, double* xbuf, int64_t N, cudaStream_t st, const int* cnt, const int* idx, int* n_general) {
    constexpr int BS = MM_CRYSTAL_BS, FBS = MM_CRYSTAL_FINISH_BS;
    constexpr size_t smem = (size_t)(MM_CRYSTAL_SCRATCH * BS + MM_CRYSTAL_TAB) * sizeof(double), fsmem = (size_t)MM_CRYSTAL_SCRATCH * FBS * sizeof(double);
    constexpr size_t smem_aset = (size_t)((HASQ ? MM_CRYSTAL_SCRATCH_ASET_Q : MM_CRYSTAL_SCRATCH_ASET) * BS + MM_CRYSTAL_TAB) * sizeof(double);      // (only when AS == 2)
    // "crystal_variant" (mm_set_tuning / MM_CRYSTAL_VARIANT): 0 = dense branch-free pass, 1 = sparse per-lane pass, 2 = phased pass.
    // Measured on configs[3] (1e7 points, step B, profiles/r02_crystal_variants.jsonl + r02_crystal_*_ncu.txt): dense 21.9 ms, sparse 21.9 ms,
    // phased 25.6 ms — the sparse pass does HALF the thread-level work of the dense one but leaves 15 of 32 lanes active per instruction, the
    // phased pass (unrolled, constant-bank operands) outgrows the 32 KB instruction cache (23 % no_instruction stalls).  Tests run all of them.
    // 3 / 4 = sparse / dense pass over chunks re-ordered by first-pass candidate count (crystal_presort_kernel); 5 = active-set pass over re-ordered
    // chunks, 6 = active-set pass in the natural order (18.3 ms; 5: 19.6 ms on SoA arrays, profiles/r02_crystal_aset_*).
    // default: the active-set pass, over re-ordered chunks for AoS arrays (a point's rows are contiguous there, so the re-ordered refills
    // stay coalesced: 18.0 against 19.3 ms) and in the natural order for SoA arrays (17.0 against 18.1 ms; profiles/r02_crystal_aset_ab_*.jsonl);
    // with cross hardening 23.1 against 27.6 ms for the dense pass (profiles/r02_crystal_family_final.log)
    const int variant = crystal_knobs().variant < 0 ? (LAYOUT == 1 ? 5 : 6) : crystal_knobs().variant;
    constexpr int SP = 1, AS = 2;
    const int pass = (variant == 1 || variant == 3) ? 1 : (((variant == 5 || variant == 6) && N < ((int64_t)1 << 31)) ? AS : 0);       // (the hand-over list holds 32-bit indices)
    const bool presort = (variant == 3 || variant == 4 || variant == 5) && !cnt && !idx;
    static SmemOptIn optin_newton, optin_newton_sparse, optin_newton_aset, optin_finish;
    if (int rc = ensure_smem(optin_newton, crystal_newton_kernel<HASQ, INTPOW, LAYOUT, BS, 0>, smem, "mm_crystal")) return rc;
    if (int rc = ensure_smem(optin_newton_sparse, crystal_newton_kernel<HASQ, INTPOW, LAYOUT, BS, SP>, smem, "mm_crystal")) return rc;
    if (int rc = ensure_smem(optin_newton_aset, crystal_newton_kernel<HASQ, INTPOW, LAYOUT, BS, AS>, AS == 2 ? smem_aset : smem, "mm_crystal")) return rc;
    if (int rc = ensure_smem(optin_finish, crystal_finish_kernel<HASQ, INTPOW, LAYOUT, FBS>, fsmem, "mm_crystal")) return rc;
    // six blocks of scratch columns need 210 KB of the SM's 228 KB: ask for the largest shared-memory carve-out (with the default heuristic
    // ncu reported a shared-memory limit of 5 blocks, i.e. 10 instead of 12 resident warps)
    static bool carve[64] = {};
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !carve[dev]) {
            cudaFuncSetAttribute(crystal_newton_kernel<HASQ, INTPOW, LAYOUT, BS, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(crystal_newton_kernel<HASQ, INTPOW, LAYOUT, BS, SP>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(crystal_newton_kernel<HASQ, INTPOW, LAYOUT, BS, AS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(crystal_finish_kernel<HASQ, INTPOW, LAYOUT, FBS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaGetLastError();
            carve[dev] = true;
        }
    }
    const int rw8 = sm_count() * MM_CRYSTAL_MINBLOCKS * (BS / 32) * 8;
    int64_t chunk = crystal_chunk_for(N, rw8);
    const CrystalKnobs& kn = crystal_knobs();
    const int thresh = kn.thresh > 0 ? kn.thresh : MM_CRYSTAL_THRESH;
    if (kn.chunk > 0) chunk = kn.chunk;
    const int64_t warps = cnt ? (N + 63) / 64 : (N + chunk - 1) / chunk;      // device-side count: cover the smallest chunk, surplus warps exit at once
    const int64_t blocks = (warps + (BS / 32) - 1) / (BS / 32);
    // workspace of the scheduling data (stream-ordered, freed below): order[N] u16, hc[chunks] u16, heavy_idx[N] i32, heavy_cnt i32
    uint16_t* order = nullptr; uint16_t* hc = nullptr; int* heavy_idx = nullptr; int* heavy_cnt = nullptr;
    char* ws = nullptr;
    const bool split = (pass == 2);                                 // active-set launch + sparse launch over the points it hands over (also for the wrappers' batches)
    const bool sorted = presort && chunk <= 1024 && N > 0;
    const int64_t nchunks = (N + chunk - 1) / chunk;
    if (sorted || split) {
        const size_t o_order = 0, o_hc = o_order + (((size_t)N * 2 + 15) & ~(size_t)15), o_idx = o_hc + (((size_t)nchunks * 2 + 15) & ~(size_t)15),
                     o_cnt = o_idx + (split ? (size_t)N * 4 : 0), total = o_cnt + 16;
        cudaError_t e = ws_alloc_async((void**)&ws, total, st);
        if (e != cudaSuccess) return set_error(MM_ERR_CUDA, "mm_crystal: cudaMallocAsync: %s", cudaGetErrorString(e));
        if (sorted) order = (uint16_t*)(ws + o_order);
        if (sorted && split) hc = (uint16_t*)(ws + o_hc);
        if (split) { heavy_idx = (int*)(ws + o_idx); heavy_cnt = (int*)(ws + o_cnt); cudaMemsetAsync(heavy_cnt, 0, sizeof(int), st); }
    }
    if (sorted) {                                                    // chunk-local re-ordering by first-pass candidate count (scheduling only)
        crystal_presort_kernel<HASQ, LAYOUT><<<(unsigned)nchunks, MM_PRESORT_BS, 0, st>>>(k, ptrs, order, N, (int)chunk, hc, heavy_idx, heavy_cnt);
        if (int rc = check_launch("mm_crystal (presort kernel)")) { cudaFreeAsync(ws, st); return rc; }
    }
    const CrystalSched sched{order, hc, heavy_idx, heavy_cnt, 0, n_general, nullptr};
    if (INTPOW && !HASQ && variant == 2) {          // "crystal_variant" 2: the phased pass (integer m, q = 0) — measured slower, kept for A/B
        constexpr size_t psmem = (size_t)MM_CRYSTAL_SCRATCH_PHASED * BS * sizeof(double);
        static SmemOptIn optin_phased;
        if (int rc = ensure_smem(optin_phased, crystal_newton_phased_kernel<LAYOUT, BS>, psmem, "mm_crystal")) return rc;
        static bool carve_p[64] = {};
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !carve_p[dev]) {
            cudaFuncSetAttribute(crystal_newton_phased_kernel<LAYOUT, BS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaGetLastError();
            carve_p[dev] = true;
        }
        CrystalNK nk;
        make_crystal_nk(nk, k);
        crystal_newton_phased_kernel<LAYOUT, BS><<<(unsigned)blocks, BS, psmem, st>>>(nk, ptrs, active, iters, xbuf, N, (int)chunk, thresh, cnt, idx, rw8, n_general);
    } else if (pass == 2) crystal_newton_kernel<HASQ, INTPOW, LAYOUT, BS, AS><<<(unsigned)blocks, BS, AS == 2 ? smem_aset : smem, st>>>(k, ptrs, active, iters, xbuf, N, (int)chunk, thresh, cnt, idx, rw8, sched);
    else if (pass == 1)   crystal_newton_kernel<HASQ, INTPOW, LAYOUT, BS, SP><<<(unsigned)blocks, BS, smem, st>>>(k, ptrs, active, iters, xbuf, N, (int)chunk, thresh, cnt, idx, rw8, sched);
    else                  crystal_newton_kernel<HASQ, INTPOW, LAYOUT, BS, 0><<<(unsigned)blocks, BS, smem, st>>>(k, ptrs, active, iters, xbuf, N, (int)chunk, thresh, cnt, idx, rw8, sched);
    int rc = check_launch("mm_crystal (Newton kernel)");
    if (rc == 0 && split) {
        // the points the active-set launch did not take (heavy from the start, or grown past MM_ASET_MAX candidates): the sparse pass over the
        // device-side list, one resident wave of warps walking its chunks, results written at each point's own index
        const CrystalSched sc2{nullptr, nullptr, nullptr, nullptr, 1, n_general, idx};      // a batch: the heavy list holds batch positions, idx[position] = where the inputs are
        const int64_t wave = (int64_t)sm_count() * MM_CRYSTAL_MINBLOCKS;
        crystal_newton_kernel<HASQ, INTPOW, LAYOUT, BS, SP><<<(unsigned)(wave < blocks ? wave : blocks), BS, smem, st>>>(k, ptrs, active, iters, xbuf, N, (int)chunk, thresh, heavy_cnt, heavy_idx, rw8, sc2);
        rc = check_launch("mm_crystal (Newton kernel, heavy points)");
    }
    if (ws) cudaFreeAsync(ws, st);
    if (rc) return rc;
    crystal_finish_kernel<HASQ, INTPOW, LAYOUT, FBS><<<(unsigned)((N + FBS - 1) / FBS), FBS, fsmem, st>>>(k, ptrs, active, n_fail, xbuf, N, cnt, idx, n_general, crystal_knobs().finish_pf);
    return check_launch("mm_crystal (finish kernel)");
}

template <bool HASQ, int LAYOUT>
int launch_crystal_by_exponent(const CrystalK& k, uint16_t* active, uint16_t* iters, int32_t* n_fail, const Ptrs<2, 3>& ptrs, double* xbuf, int64_t N, cudaStream_t st, const int* cnt, const int* idx, int* n_general) {
    return k.m_int ? launch_crystal_two_kernel<HASQ, true, LAYOUT>(k, active, iters, n_fail, ptrs, xbuf, N, st, cnt, idx, n_general)
                   : launch_crystal_two_kernel<HASQ, false, LAYOUT>(k, active, iters, n_fail, ptrs, xbuf, N, st, cnt, idx, n_general);
}

int launch_crystal_general(const CrystalK& k, uint16_t* active, uint16_t* iters, int32_t* n_fail, const Ptrs<2, 3>& ptrs, int64_t N, int layout, cudaStream_t st) {
    if (N == 0) return MM_OK;
    CrystalGeneralOp op;
    op.k = k; op.active = active; op.iters = iters; op.n_fail = n_fail;
    return launch_points<CrystalGeneralOp, 64>(op, ptrs, N, layout, st, "mm_crystal");
}

template <bool HASQ>
int launch_crystal_structured(const CrystalK& k, uint16_t* active, uint16_t* iters, int32_t* n_fail, const Ptrs<2, 3>& ptrs, double* xbuf, int64_t N, int layout, cudaStream_t st,
                              const int* cnt = nullptr, const int* idx = nullptr) {
    if (N == 0) return MM_OK;
    if (layout != 0 && layout != 1) return set_error(MM_ERR_ARG, "mm_crystal: unknown layout %d", layout);
    // how many points the Newton / finish kernels left to the fix-up pass: none on the benchmark configuration, and then the fix-up blocks exit at
    // once instead of scanning the flag array (0.21 ms per 1e7 points)
    int* n_general = nullptr;
    {
        cudaError_t e = ws_alloc_async((void**)&n_general, 16, st);
        if (e != cudaSuccess) return set_error(MM_ERR_CUDA, "mm_crystal: cudaMallocAsync: %s", cudaGetErrorString(e));
        cudaMemsetAsync(n_general, 0, sizeof(int), st);
    }
    int rc;
    if (layout == 0) rc = launch_crystal_by_exponent<HASQ, 0>(k, active, iters, n_fail, ptrs, xbuf, N, st, cnt, idx, n_general);
    else             rc = launch_crystal_by_exponent<HASQ, 1>(k, active, iters, n_fail, ptrs, xbuf, N, st, cnt, idx, n_general);
    if (rc) { cudaFreeAsync(n_general, st); return rc; }
    const int64_t nranges = (N + MM_FIXUP_RANGE - 1) / MM_FIXUP_RANGE;
    const int per_sm = crystal_knobs().fixup_per_sm > 0 ? crystal_knobs().fixup_per_sm : 5;
    int64_t blocks = (int64_t)sm_count() * per_sm;
    if (blocks > nranges) blocks = nranges;
    if (layout == 0) crystal_fixup_kernel<0><<<(unsigned)blocks, MM_FIXUP_BS, 0, st>>>(k, ptrs, active, iters, n_fail, N, nranges, cnt, idx, n_general);
    else             crystal_fixup_kernel<1><<<(unsigned)blocks, MM_FIXUP_BS, 0, st>>>(k, ptrs, active, iters, n_fail, N, nranges, cnt, idx, n_general);
    cudaFreeAsync(n_general, st);
    return check_launch("mm_crystal (fix-up pass)");
}

// The structured two-kernel path on a BATCH of points of SoA arrays, for the dimension wrappers (mm_wrap.cu): idx[0 .. *cnt) lists the
// points to update (device-side list and count).  The INPUTS (strain increment, old state) are read where the point is, in[a][c*N + idx[e]];
// the OUTPUTS (σ, ∂σ∂ε, new state, active, iters) are packed by list position e.  N = stride of every array.  Requires k.h_struct.
int crystal_structured_batch(const CrystalK& k, uint16_t* active, uint16_t* iters, const Ptrs<2, 3>& ptrs, int64_t N, const int* cnt, const int* idx, cudaStream_t st) {
    return (k.b_h != 0.0) ? launch_crystal_structured<true>(k, active, iters, nullptr, ptrs, nullptr, N, 0, st, cnt, idx)
                          : launch_crystal_structured<false>(k, active, iters, nullptr, ptrs, nullptr, N, 0, st, cnt, idx);
}
bool crystal_force_general() { return crystal_knobs().force_general != 0; }

}  // namespace mm

using namespace mm;

extern "C" int mm_crystal(const MMCrystal* p, const double* deps, double dt, const double* state_in, double* sig, double* tan,
                          double* state_out, uint16_t* active, uint16_t* iters, int32_t* n_fail, const MMNewtonOpts* opts,
                          int64_t N, int layout, void* stream) {
    if (!p || N < 0 || (N > 0 && (!deps || !state_in || !sig || !state_out))) return set_error(MM_ERR_ARG, "mm_crystal: bad argument");
    if (p->nslip != MM_NSLIP) return set_error(MM_ERR_ARG, "mm_crystal: nslip must be %d (got %d)", MM_NSLIP, p->nslip);
    CrystalK k;
    make_crystal_k(k, p->E, p->nu, p->tau_y, p->H_kin, p->alpha_inf, p->t_star, p->sigma_c, p->m, p->MS, p->H, dt,
                   opts ? opts->tol : 1e-8, opts ? opts->max_iter : 100);
    Ptrs<2, 3> ptrs; ptrs.in[0] = deps; ptrs.in[1] = state_in; ptrs.out[0] = sig; ptrs.out[1] = tan; ptrs.out[2] = state_out;
    cudaStream_t st = (cudaStream_t)stream;
    if (!k.h_struct || crystal_knobs().force_general) return launch_crystal_general(k, active, iters, n_fail, ptrs, N, layout, st);
    // the structured path needs a per-point flag array for its fix-up pass; callers normally pass `active`,
    // otherwise a stream-ordered temporary is used (the only allocation this library ever makes on a compute call)
    uint16_t* flags = active;
    if (!flags && N > 0) {
        cudaError_t e = ws_alloc_async((void**)&flags, (size_t)N * sizeof(uint16_t), st);
        if (e != cudaSuccess) return set_error(MM_ERR_CUDA, "mm_crystal: cudaMallocAsync: %s", cudaGetErrorString(e));
    }
    double* xbuf = nullptr;
    if (state_in == state_out && N > 0) {        // in-place update: see DirectAccRW
        cudaError_t e = ws_alloc_async((void**)&xbuf, (size_t)N * 12 * sizeof(double), st);
        if (e != cudaSuccess) { if (!active && flags) cudaFreeAsync(flags, st); return set_error(MM_ERR_CUDA, "mm_crystal: cudaMallocAsync: %s", cudaGetErrorString(e)); }
    }
    const int rc = (k.b_h != 0.0) ? launch_crystal_structured<true>(k, flags, iters, n_fail, ptrs, xbuf, N, layout, st)
                                  : launch_crystal_structured<false>(k, flags, iters, n_fail, ptrs, xbuf, N, layout, st);
    if (xbuf) cudaFreeAsync(xbuf, st);
    if (!active && flags) cudaFreeAsync(flags, st);
    return rc;
}

// test / tuning hook (not part of the reference-facing surface): returns the previous value, -1 for an unknown name
extern "C" int mm_set_tuning(const char* name, int value) {
    if (!name) return -1;
    CrystalKnobs& k = crystal_knobs();
    int* slot = nullptr;
    if (!strcmp(name, "crystal_force_general")) slot = &k.force_general;
    else if (!strcmp(name, "crystal_thresh")) slot = &k.thresh;
    else if (!strcmp(name, "crystal_chunk")) slot = &k.chunk;
    else if (!strcmp(name, "crystal_fixup_per_sm")) slot = &k.fixup_per_sm;
    else if (!strcmp(name, "crystal_variant")) slot = &k.variant;
    else if (!strcmp(name, "crystal_finish_prefetch")) slot = &k.finish_pf;
    if (!slot) slot = mm::wrap_tuning_slot(name);              // "wrapped_j2_*" knobs live next to their kernel (mm_wrap.cu)
    if (!slot) slot = mm::ops_tuning_slot(name);               // "xu_soa_tiled" (mm_ops.cu)
    if (!slot) slot = mm::host_tuning_slot(name);              // "mmh_plastic_via_factored" (mm_host.cu)
    if (!slot) return -1;
    const int old = *slot;
    *slot = value;
    return old;
}
