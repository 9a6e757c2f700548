// Library plumbing (errors, device info) and the host-side parameter helpers that mirror the
// reference's Julia constructors (run once per material; nothing here is on the per-point path).
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include "../../include/mmb200.h"

namespace mm {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(MM_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    }
    return MM_OK;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// Stream-ordered workspace for the few entry points that need one (the crystal fix-up flags / aliased-state side buffer, the rounds of
// mm_wrapped_crystal): the library's OWN memory pool per device with an unlimited release threshold, so a workspace freed by one call is
// reused by the next instead of going back to the driver at every synchronisation (with the device's default pool and its zero
// threshold a 2.3 GB workspace cost 12 ms of cuMemMap per call — half the call, profiles/r02_wrapped_crystal_rounds.log).
cudaError_t ws_alloc_async(void** ptr, size_t bytes, cudaStream_t st) {
    static cudaMemPool_t pools[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaMallocAsync(ptr, bytes, st);
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);                    // pool creation only; the allocation itself is a stream-ordered call
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool;
        e = cudaMemPoolCreate(&pool, &props);
        if (e != cudaSuccess) { cudaGetLastError(); return cudaMallocAsync(ptr, bytes, st); }
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        pools[dev] = pool;
    }
    return cudaMallocFromPoolAsync(ptr, bytes, pools[dev], st);
}

}  // namespace mm

extern "C" {

const char* mm_version(void) { return "mmb200 0.1.0 (sm_100a)"; }
const char* mm_last_error(void) { return mm::g_err; }

int mm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// elastic_tangent_3D(E, ν) — reference src/LinearElastic.jl:27-35.  C_ijkl = λ δij δkl + μ (δik δjl + δil δjk),
// stored as SymmetricTensor{4,3}: out36[I + 6 J].
int mm_elastic_tangent_3d(double E, double nu, double* out36) {
    if (!out36) return mm::set_error(MM_ERR_ARG, "mm_elastic_tangent_3d: null output");
    const double lam = E * nu / ((1 + nu) * (1 - 2 * nu));
    const double mu = E / (2 * (1 + nu));
    static const int pi[6] = {0, 1, 2, 1, 2, 2}, pj[6] = {0, 0, 0, 1, 1, 2};
    for (int J = 0; J < 6; ++J)
        for (int I = 0; I < 6; ++I) {
            const int i = pi[I], j = pj[I], k = pi[J], l = pj[J];
            out36[I + 6 * J] = lam * (i == j) * (k == l) + mu * ((i == k) * (j == l) + (i == l) * (j == k));
        }
    return MM_OK;
}

// slipsystems(::FCC/::BCC, R::RodriguesParam) — reference src/CrystalViscoPlastic/slipsystems.jl:9-53.
// Plane-major / direction-minor order; every plane normal and direction rotated by R then normalised.
// RodriguesParam(x,y,z) is the rotation of the unit quaternion (1,x,y,z)/sqrt(1+x²+y²+z²).
int mm_slipsystems(int structure, const double rodrigues[3], double* planes, double* dirs) {
    if (!rodrigues || !planes || !dirs) return mm::set_error(MM_ERR_ARG, "mm_slipsystems: null argument");
    static const double fccP[4][3] = {{1, 1, 1}, {1, 1, -1}, {1, -1, -1}, {1, -1, 1}};
    static const double fccD[12][3] = {{1, 0, -1}, {0, 1, -1}, {1, -1, 0}, {0, 1, 1}, {1, 0, 1}, {1, -1, 0},
                                       {1, 0, 1},  {0, 1, -1}, {1, 1, 0},  {0, 1, 1}, {1, 0, -1}, {1, 1, 0}};
    static const double bccP[6][3] = {{1, 1, 0}, {1, -1, 0}, {1, 0, 1}, {-1, 0, -1}, {0, 1, 1}, {0, 1, -1}};
    static const double bccD[12][3] = {{1, -1, 1}, {1, -1, -1}, {1, 1, 1},  {1, 1, -1},  {-1, 1, 1}, {-1, -1, 1},
                                       {1, 1, 1},  {1, -1, 1},  {1, 1, -1}, {-1, 1, -1}, {-1, 1, 1}, {1, 1, 1}};
    if (structure != MM_FCC && structure != MM_BCC) return mm::set_error(MM_ERR_ARG, "mm_slipsystems: unknown structure %d", structure);
    const double x = rodrigues[0], y = rodrigues[1], z = rodrigues[2];
    const double s = 1.0 / sqrt(1.0 + x * x + y * y + z * z);
    const double qw = s, qx = x * s, qy = y * s, qz = z * s;
    const double R[3][3] = {{1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qw * qz), 2 * (qx * qz + qw * qy)},
                            {2 * (qx * qy + qw * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qw * qx)},
                            {2 * (qx * qz - qw * qy), 2 * (qy * qz + qw * qx), 1 - 2 * (qx * qx + qy * qy)}};
    const int per_plane = (structure == MM_FCC) ? 3 : 2;
    for (int sidx = 0; sidx < MM_NSLIP; ++sidx) {
        const double* pv = (structure == MM_FCC) ? fccP[sidx / per_plane] : bccP[sidx / per_plane];
        const double* dv = (structure == MM_FCC) ? fccD[sidx] : bccD[sidx];
        double rp[3], rd[3];
        for (int i = 0; i < 3; ++i) {
            rp[i] = R[i][0] * pv[0] + R[i][1] * pv[1] + R[i][2] * pv[2];
            rd[i] = R[i][0] * dv[0] + R[i][1] * dv[1] + R[i][2] * dv[2];
        }
        const double np = sqrt(rp[0] * rp[0] + rp[1] * rp[1] + rp[2] * rp[2]);
        const double nd = sqrt(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
        for (int i = 0; i < 3; ++i) { planes[sidx * 3 + i] = rp[i] / np; dirs[sidx * 3 + i] = rd[i] / nd; }
    }
    return MM_NSLIP;
}

// CrystalViscoPlastic(...) constructor — reference CrystalViscoPlastic.jl:38-45:
// H = H_iso (q ones + (1-q) I),  MS[i] = m_i ⊗ s_i.
int mm_crystal_setup(MMCrystal* p, const double* planes, const double* dirs) {
    if (!p || !planes || !dirs) return mm::set_error(MM_ERR_ARG, "mm_crystal_setup: null argument");
    p->nslip = MM_NSLIP;
    p->pad_ = 0;
    for (int i = 0; i < MM_NSLIP; ++i)
        for (int j = 0; j < MM_NSLIP; ++j) p->H[i][j] = p->H_iso * (p->q + (1 - p->q) * (i == j ? 1.0 : 0.0));
    for (int s = 0; s < MM_NSLIP; ++s)
        for (int col = 0; col < 3; ++col)
            for (int row = 0; row < 3; ++row) p->MS[s][row + 3 * col] = planes[s * 3 + row] * dirs[s * 3 + col];
    return MM_OK;
}

// XuNeedleman(; σₘₐₓ, τₘₐₓ, Φₙ, Φₜ, Δₙˢ) — reference XuNeedleman.jl:30-34
int mm_xu_needleman_setup(double sigma_max, double tau_max, double Phi_n, double Phi_t, double Delta_n_star, MMXuNeedleman* out) {
    if (!out) return mm::set_error(MM_ERR_ARG, "mm_xu_needleman_setup: null output");
    out->Phi_n = Phi_n;
    out->Phi_t = Phi_t;
    out->delta_n = Phi_n / (exp(1.0) * sigma_max);
    out->delta_t = Phi_t / (sqrt(0.5 * exp(1.0)) * tau_max);
    out->Delta_n_star = Delta_n_star;
    return MM_OK;
}

}  // extern "C"
