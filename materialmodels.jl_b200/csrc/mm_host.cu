// mmh_* — host-buffer entry points: the same kernels behind a pipelined H2D -> kernel -> D2H chunk
// loop.  This is the call a host-resident FE assembler makes; PCIe, not the kernel, bounds it, so the
// only design goals are (i) keep both copy engines and the SMs busy at once (NSLOT streams, one chunk
// in flight per slot) and (ii) never allocate per call (workspace grows once and is reused).
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "../../include/mmb200.h"

namespace mm {
int set_error(int code, const char* fmt, ...);

namespace {
constexpr int NSLOT = 3;
int64_t g_chunk = 1 << 20;

struct Arr {
    const void* h_in;   // host source (inputs) or NULL
    void* h_out;        // host destination (outputs) or NULL
    int nc;             // components per point
    int elem;           // bytes per component
    bool per_layout;    // false: one value per point, layout-independent (flags / iters / active)
    void* d[NSLOT];     // device chunk buffers
};

struct Workspace {
    void* base = nullptr;
    size_t bytes = 0;
    cudaStream_t streams[NSLOT] = {nullptr, nullptr, nullptr};
    int32_t* d_fail = nullptr;
    bool init = false;
};
Workspace g_ws_dev[64];             // one workspace (streams, staging buffer, counter) per device the process has used
int g_cur = 0;
#define g_ws g_ws_dev[g_cur]

int ensure(size_t bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_error(MM_ERR_NOGPU, "mmh: no CUDA device");
    g_cur = (dev >= 0 && dev < 64) ? dev : 0;
    if (!g_ws.init) {
        for (int s = 0; s < NSLOT; ++s)
            if (cudaStreamCreateWithFlags(&g_ws.streams[s], cudaStreamNonBlocking) != cudaSuccess) return set_error(MM_ERR_CUDA, "mmh: cudaStreamCreate failed");
        if (cudaMalloc(&g_ws.d_fail, sizeof(int32_t)) != cudaSuccess) return set_error(MM_ERR_CUDA, "mmh: cudaMalloc failed");
        g_ws.init = true;
    }
    if (bytes > g_ws.bytes) {
        if (g_ws.base) cudaFree(g_ws.base);
        g_ws.base = nullptr; g_ws.bytes = 0;
        cudaError_t e = cudaMalloc(&g_ws.base, bytes);
        if (e != cudaSuccess) return set_error(MM_ERR_CUDA, "mmh: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        g_ws.bytes = bytes;
    }
    return MM_OK;
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// copy points [i0, i0+n) of a host array <-> dense device chunk (chunk seen by the kernel as an array of n points)
int copy_chunk(const Arr& a, int slot, int64_t N, int64_t i0, int64_t n, int layout, bool to_device, cudaStream_t st) {
    cudaError_t e;
    char* dev = (char*)a.d[slot];
    if (layout == MM_LAYOUT_SOA && a.per_layout && a.nc > 1) {
        const size_t width = (size_t)n * a.elem, hpitch = (size_t)N * a.elem;
        if (to_device) e = cudaMemcpy2DAsync(dev, width, (const char*)a.h_in + (size_t)i0 * a.elem, hpitch, width, a.nc, cudaMemcpyHostToDevice, st);
        else           e = cudaMemcpy2DAsync((char*)a.h_out + (size_t)i0 * a.elem, hpitch, dev, width, width, a.nc, cudaMemcpyDeviceToHost, st);
    } else {
        const size_t off = (size_t)i0 * a.nc * a.elem, len = (size_t)n * a.nc * a.elem;
        if (to_device) e = cudaMemcpyAsync(dev, (const char*)a.h_in + off, len, cudaMemcpyHostToDevice, st);
        else           e = cudaMemcpyAsync((char*)a.h_out + off, dev, len, cudaMemcpyDeviceToHost, st);
    }
    if (e != cudaSuccess) return set_error(MM_ERR_CUDA, "mmh: memcpy failed: %s", cudaGetErrorString(e));
    return MM_OK;
}

// launch(slot, n, stream) must enqueue the device-pointer entry point on the slot's chunk buffers
template <class Launch>
int pipeline(Arr* arrs, int narr, int64_t N, int layout, int32_t* n_fail, Launch launch) {
    if (N == 0) return MM_OK;
    static std::mutex mu;                               // the staging workspace is shared: host threads take turns
    std::lock_guard<std::mutex> lock(mu);
    const int64_t chunk = (N < g_chunk) ? N : g_chunk;
    size_t per_slot = 0;
    for (int a = 0; a < narr; ++a) if (arrs[a].h_in || arrs[a].h_out) per_slot += align256((size_t)chunk * arrs[a].nc * arrs[a].elem);
    int rc = ensure(per_slot * NSLOT);
    if (rc) return rc;
    char* base = (char*)g_ws.base;
    for (int s = 0; s < NSLOT; ++s) {
        size_t off = 0;
        for (int a = 0; a < narr; ++a) {
            if (arrs[a].h_in || arrs[a].h_out) { arrs[a].d[s] = base + s * per_slot + off; off += align256((size_t)chunk * arrs[a].nc * arrs[a].elem); }
            else arrs[a].d[s] = nullptr;
        }
    }
    if (cudaMemsetAsync(g_ws.d_fail, 0, sizeof(int32_t), g_ws.streams[0]) != cudaSuccess) return set_error(MM_ERR_CUDA, "mmh: memset failed");
    if (cudaStreamSynchronize(g_ws.streams[0]) != cudaSuccess) return set_error(MM_ERR_CUDA, "mmh: sync failed");
    int64_t k = 0;
    for (int64_t i0 = 0; i0 < N; i0 += chunk, ++k) {
        const int slot = (int)(k % NSLOT);
        const int64_t n = (N - i0 < chunk) ? (N - i0) : chunk;
        cudaStream_t st = g_ws.streams[slot];
        for (int a = 0; a < narr; ++a)
            if (arrs[a].h_in) { rc = copy_chunk(arrs[a], slot, N, i0, n, layout, true, st); if (rc) return rc; }
        rc = launch(slot, n, st);
        if (rc) return rc;
        for (int a = 0; a < narr; ++a)
            if (arrs[a].h_out) { rc = copy_chunk(arrs[a], slot, N, i0, n, layout, false, st); if (rc) return rc; }
    }
    for (int s = 0; s < NSLOT; ++s)
        if (cudaStreamSynchronize(g_ws.streams[s]) != cudaSuccess) return set_error(MM_ERR_CUDA, "mmh: stream sync failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (n_fail) {
        int32_t f = 0;
        if (cudaMemcpy(&f, g_ws.d_fail, sizeof(f), cudaMemcpyDeviceToHost) != cudaSuccess) return set_error(MM_ERR_CUDA, "mmh: n_fail readback failed");
        *n_fail += f;
    }
    return MM_OK;
}

inline Arr in_arr(const void* h, int nc, int elem = 8) { Arr a; memset(&a, 0, sizeof(a)); a.h_in = h; a.nc = nc; a.elem = elem; a.per_layout = true; return a; }
inline Arr out_arr(void* h, int nc, int elem = 8, bool per_layout = true) { Arr a; memset(&a, 0, sizeof(a)); a.h_out = h; a.nc = nc; a.elem = elem; a.per_layout = per_layout; return a; }
}  // namespace
}  // namespace mm

using namespace mm;

extern "C" {

void* mmh_alloc_pinned(uint64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); set_error(MM_ERR_CUDA, "mmh_alloc_pinned(%llu) failed", (unsigned long long)bytes); return nullptr; }
    return p;
}
void mmh_free_pinned(void* p) { if (p) cudaFreeHost(p); }
int64_t mmh_set_chunk_points(int64_t n) { int64_t old = g_chunk; if (n > 0) g_chunk = n; return old; }

int mmh_linear_elastic(const MMLinearElastic* p, const double* eps, double* sig, double* tan_or_null, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!eps || !sig))) return set_error(MM_ERR_ARG, "mmh_linear_elastic: bad argument");
    Arr a[3] = {in_arr(eps, 6), out_arr(sig, 6), out_arr(tan_or_null, 36)};
    return pipeline(a, 3, N, layout, nullptr, [&](int s, int64_t n, cudaStream_t st) {
        return mm_linear_elastic(p, (const double*)a[0].d[s], (double*)a[1].d[s], (double*)a[2].d[s], n, layout, st);
    });
}

int mmh_plastic(const MMPlastic* p, const double* eps, const double* state_in, double* sig, double* tan, double* state_out,
                uint8_t* flags, uint8_t* iters, int32_t* n_fail, const MMNewtonOpts* opts, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!eps || !state_in || !sig || !state_out))) return set_error(MM_ERR_ARG, "mmh_plastic: bad argument");
    Arr a[7] = {in_arr(eps, 6), in_arr(state_in, 14), out_arr(sig, 6), out_arr(tan, 36), out_arr(state_out, 14),
                out_arr(flags, 1, 1, false), out_arr(iters, 1, 1, false)};
    return pipeline(a, 7, N, layout, n_fail, [&](int s, int64_t n, cudaStream_t st) {
        return mm_plastic(p, (const double*)a[0].d[s], (const double*)a[1].d[s], (double*)a[2].d[s], (double*)a[3].d[s], (double*)a[4].d[s],
                          (uint8_t*)a[5].d[s], (uint8_t*)a[6].d[s], g_ws.d_fail, opts, n, layout, st);
    });
}

int mmh_hyper(int model, const MMHyper* p, const double* strain, int strain_kind, double* S, double* dSdC, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!strain || !S))) return set_error(MM_ERR_ARG, "mmh_hyper: bad argument");
    Arr a[3] = {in_arr(strain, strain_kind == MM_STRAIN_F ? 9 : 6), out_arr(S, 6), out_arr(dSdC, 36)};
    return pipeline(a, 3, N, layout, nullptr, [&](int s, int64_t n, cudaStream_t st) {
        return mm_hyper(model, p, (const double*)a[0].d[s], strain_kind, (double*)a[1].d[s], (double*)a[2].d[s], n, layout, st);
    });
}

int mmh_crystal(const MMCrystal* p, const double* deps, double dt, const double* state_in, double* sig, double* tan, double* state_out,
                uint16_t* active, uint8_t* iters, int32_t* n_fail, const MMNewtonOpts* opts, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!deps || !state_in || !sig || !state_out))) return set_error(MM_ERR_ARG, "mmh_crystal: bad argument");
    Arr a[7] = {in_arr(deps, 6), in_arr(state_in, 42), out_arr(sig, 6), out_arr(tan, 36), out_arr(state_out, 42),
                out_arr(active, 1, 2, false), out_arr(iters, 1, 1, false)};
    return pipeline(a, 7, N, layout, n_fail, [&](int s, int64_t n, cudaStream_t st) {
        return mm_crystal(p, (const double*)a[0].d[s], dt, (const double*)a[1].d[s], (double*)a[2].d[s], (double*)a[3].d[s], (double*)a[4].d[s],
                          (uint16_t*)a[5].d[s], (uint8_t*)a[6].d[s], g_ws.d_fail, opts, n, layout, st);
    });
}

int mmh_xu_needleman(const MMXuNeedleman* p, const double* jump, double* T, double* dTdD, int dim, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!jump || !T)) || (dim != 2 && dim != 3)) return set_error(MM_ERR_ARG, "mmh_xu_needleman: bad argument");
    Arr a[3] = {in_arr(jump, dim), out_arr(T, dim), out_arr(dTdD, dim * dim)};
    return pipeline(a, 3, N, layout, nullptr, [&](int s, int64_t n, cudaStream_t st) {
        return mm_xu_needleman(p, (const double*)a[0].d[s], (double*)a[1].d[s], (double*)a[2].d[s], dim, n, layout, st);
    });
}

int mmh_wrapped_linear_elastic(int dim_kind, const MMLinearElastic* p, const double* eps_red, double* sig_red, double* tan_red, uint8_t* outer_iters,
                               int32_t* n_fail, const MMNewtonOpts* wrapper_opts, int64_t N, int layout) {
    const int nc = mm_wrap_ncomp(dim_kind);
    if (nc < 0) return nc;
    if (!p || N < 0 || (N > 0 && (!eps_red || !sig_red))) return set_error(MM_ERR_ARG, "mmh_wrapped_linear_elastic: bad argument");
    Arr a[4] = {in_arr(eps_red, nc), out_arr(sig_red, nc), out_arr(tan_red, nc * nc), out_arr(outer_iters, 1, 1, false)};
    return pipeline(a, 4, N, layout, n_fail, [&](int s, int64_t n, cudaStream_t st) {
        return mm_wrapped_linear_elastic(dim_kind, p, (const double*)a[0].d[s], (double*)a[1].d[s], (double*)a[2].d[s], (uint8_t*)a[3].d[s], g_ws.d_fail,
                                         wrapper_opts, n, layout, st);
    });
}

int mmh_wrapped_plastic(int dim_kind, const MMPlastic* p, const double* eps_red, const double* state_in, double* sig_red, double* tan_red, double* state_out,
                        uint8_t* flags, uint8_t* iters, uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* opts, const MMNewtonOpts* wrapper_opts,
                        int64_t N, int layout) {
    const int nc = mm_wrap_ncomp(dim_kind);
    if (nc < 0) return nc;
    if (!p || N < 0 || (N > 0 && (!eps_red || !state_in || !sig_red || !state_out))) return set_error(MM_ERR_ARG, "mmh_wrapped_plastic: bad argument");
    Arr a[8] = {in_arr(eps_red, nc), in_arr(state_in, 14), out_arr(sig_red, nc), out_arr(tan_red, nc * nc), out_arr(state_out, 14),
                out_arr(flags, 1, 1, false), out_arr(iters, 1, 1, false), out_arr(outer_iters, 1, 1, false)};
    return pipeline(a, 8, N, layout, n_fail, [&](int s, int64_t n, cudaStream_t st) {
        return mm_wrapped_plastic(dim_kind, p, (const double*)a[0].d[s], (const double*)a[1].d[s], (double*)a[2].d[s], (double*)a[3].d[s], (double*)a[4].d[s],
                                  (uint8_t*)a[5].d[s], (uint8_t*)a[6].d[s], (uint8_t*)a[7].d[s], g_ws.d_fail, opts, wrapper_opts, n, layout, st);
    });
}

int mmh_wrapped_crystal(int dim_kind, const MMCrystal* p, const double* deps_red, double dt, const double* state_in, double* sig_red, double* tan_red,
                        double* state_out, uint16_t* active, uint8_t* iters, uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* opts,
                        const MMNewtonOpts* wrapper_opts, int64_t N, int layout) {
    const int nc = mm_wrap_ncomp(dim_kind);
    if (nc < 0) return nc;
    if (!p || N < 0 || (N > 0 && (!deps_red || !state_in || !sig_red || !state_out))) return set_error(MM_ERR_ARG, "mmh_wrapped_crystal: bad argument");
    Arr a[8] = {in_arr(deps_red, nc), in_arr(state_in, 42), out_arr(sig_red, nc), out_arr(tan_red, nc * nc), out_arr(state_out, 42),
                out_arr(active, 1, 2, false), out_arr(iters, 1, 1, false), out_arr(outer_iters, 1, 1, false)};
    return pipeline(a, 8, N, layout, n_fail, [&](int s, int64_t n, cudaStream_t st) {
        return mm_wrapped_crystal(dim_kind, p, (const double*)a[0].d[s], dt, (const double*)a[1].d[s], (double*)a[2].d[s], (double*)a[3].d[s],
                                  (double*)a[4].d[s], (uint16_t*)a[5].d[s], (uint8_t*)a[6].d[s], (uint8_t*)a[7].d[s], g_ws.d_fail, opts, wrapper_opts, n,
                                  layout, st);
    });
}

int mmh_wrapped_transversely_isotropic(int dim_kind, const MMTransverselyIsotropic* p, const double* eps_red, const double* a3, double* sig_red,
                                       double* tan_red, uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* wrapper_opts, int64_t N, int layout) {
    const int nc = mm_wrap_ncomp(dim_kind);
    if (nc < 0) return nc;
    if (!p || N < 0 || (N > 0 && (!eps_red || !a3 || !sig_red))) return set_error(MM_ERR_ARG, "mmh_wrapped_transversely_isotropic: bad argument");
    Arr a[5] = {in_arr(eps_red, nc), in_arr(a3, 3), out_arr(sig_red, nc), out_arr(tan_red, nc * nc), out_arr(outer_iters, 1, 1, false)};
    return pipeline(a, 5, N, layout, n_fail, [&](int s, int64_t n, cudaStream_t st) {
        return mm_wrapped_transversely_isotropic(dim_kind, p, (const double*)a[0].d[s], (const double*)a[1].d[s], (double*)a[2].d[s], (double*)a[3].d[s],
                                                 (uint8_t*)a[4].d[s], g_ws.d_fail, wrapper_opts, n, layout, st);
    });
}

int mmh_hyper_ex(int model, const MMHyper* p, const double* strain, int strain_kind, int tangent_kind, double* stress, double* tangent, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!strain || !stress))) return set_error(MM_ERR_ARG, "mmh_hyper_ex: bad argument");
    const bool pf = tangent_kind == MM_TANGENT_DPDF;
    Arr a[3] = {in_arr(strain, strain_kind == MM_STRAIN_F ? 9 : 6), out_arr(stress, pf ? 9 : 6), out_arr(tangent, pf ? 81 : 36)};
    return pipeline(a, 3, N, layout, nullptr, [&](int s, int64_t n, cudaStream_t st) {
        return mm_hyper_ex(model, p, (const double*)a[0].d[s], strain_kind, tangent_kind, (double*)a[1].d[s], (double*)a[2].d[s], n, layout, st);
    });
}

int mmh_transversely_isotropic(const MMTransverselyIsotropic* p, const double* eps, const double* a3, double* sig, double* tan, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!eps || !a3 || !sig))) return set_error(MM_ERR_ARG, "mmh_transversely_isotropic: bad argument");
    Arr a[4] = {in_arr(eps, 6), in_arr(a3, 3), out_arr(sig, 6), out_arr(tan, 36)};
    return pipeline(a, 4, N, layout, nullptr, [&](int s, int64_t n, cudaStream_t st) {
        return mm_transversely_isotropic(p, (const double*)a[0].d[s], (const double*)a[1].d[s], (double*)a[2].d[s], (double*)a[3].d[s], n, layout, st);
    });
}

int mmh_linear_elastic_energy(const MMLinearElastic* p, const double* eps, double* psi, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!eps || !psi))) return set_error(MM_ERR_ARG, "mmh_linear_elastic_energy: bad argument");
    Arr a[2] = {in_arr(eps, 6), out_arr(psi, 1, 8, false)};
    return pipeline(a, 2, N, layout, nullptr, [&](int s, int64_t n, cudaStream_t st) {
        return mm_linear_elastic_energy(p, (const double*)a[0].d[s], (double*)a[1].d[s], n, layout, st);
    });
}

int mmh_hyper_energy(int model, const MMHyper* p, const double* strain, int strain_kind, double* psi, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!strain || !psi))) return set_error(MM_ERR_ARG, "mmh_hyper_energy: bad argument");
    Arr a[2] = {in_arr(strain, strain_kind == MM_STRAIN_F ? 9 : 6), out_arr(psi, 1, 8, false)};
    return pipeline(a, 2, N, layout, nullptr, [&](int s, int64_t n, cudaStream_t st) {
        return mm_hyper_energy(model, p, (const double*)a[0].d[s], strain_kind, (double*)a[1].d[s], n, layout, st);
    });
}

}  // extern "C"
