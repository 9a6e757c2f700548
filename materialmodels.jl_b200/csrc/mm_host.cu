// mmh_* — host-buffer entry points: the same kernels behind a pipelined H2D -> kernel -> D2H chunk
// loop.  This is the call a host-resident FE assembler makes; PCIe, not the kernel, bounds it, so the
// only design goals are (i) keep both copy engines and the SMs busy at once (NSLOT streams, one chunk
// in flight per slot) and (ii) never allocate per call (workspace grows once and is reused).
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include "../../include/mmb200.h"

namespace mm {
int set_error(int code, const char* fmt, ...);
int* host_tuning_slot(const char* name);
void fill_constant_tangent(double lam, double G, double* tan, int64_t N, int layout, int T);   // mm_expand.cpp
int expand_factored_chunk(const MMPlastic* p, const uint8_t* fl, const double* rows, int64_t n_rows, double* tan, int64_t N, int64_t i0, int64_t n, int layout, int T);   // mm_expand.cpp

namespace {
constexpr int NSLOT = 4;
int64_t g_chunk = 1 << 20;
std::mutex g_mu;                    // ONE lock for every mmh_* entry point: they share the staging workspace, its streams and the fail counter

struct Arr {
    const void* h_in;   // host source (inputs) or NULL
    void* h_out;        // host destination (outputs) or NULL
    int nc;             // components per point
    int elem;           // bytes per component
    bool per_layout;    // false: one value per point, layout-independent (flags / iters / active)
    void* d_res;        // device-RESIDENT array of all N points, read and updated in place (or NULL): an AoS chunk is used where it lies,
                        // an SoA chunk is staged through the slot buffer with device-to-device 2-D copies
    void* d[NSLOT];     // device chunk buffers
};

struct Workspace {
    void* base = nullptr;
    size_t bytes = 0;
    cudaStream_t streams[NSLOT] = {};
    cudaEvent_t ev[NSLOT] = {};
    int64_t* h_counts = nullptr;    // pinned: per-slot row counts of the compact return
    int32_t* d_fail = nullptr;
    cudaEvent_t ev2[NSLOT] = {};    // dense-through-factored mode: a slot's rows and flags have landed in the pinned staging
    char* h_stage = nullptr;        // pinned staging of that mode (rows + flags per slot)
    size_t h_stage_bytes = 0;
    bool init = false;
};
Workspace g_ws_dev[64];             // one workspace (streams, staging buffer, counter) per device the process has used

// callers hold g_mu
int ensure(size_t bytes, Workspace** out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_error(MM_ERR_NOGPU, "mmh: no CUDA device");
    Workspace& ws = g_ws_dev[(dev >= 0 && dev < 64) ? dev : 0];
    if (!ws.init) {
        for (int s = 0; s < NSLOT; ++s) {
            if (cudaStreamCreateWithFlags(&ws.streams[s], cudaStreamNonBlocking) != cudaSuccess) return set_error(MM_ERR_CUDA, "mmh: cudaStreamCreate failed");
            if (cudaEventCreateWithFlags(&ws.ev[s], cudaEventDisableTiming) != cudaSuccess) return set_error(MM_ERR_CUDA, "mmh: cudaEventCreate failed");
            if (cudaEventCreateWithFlags(&ws.ev2[s], cudaEventDisableTiming) != cudaSuccess) return set_error(MM_ERR_CUDA, "mmh: cudaEventCreate failed");
        }
        if (cudaMalloc(&ws.d_fail, sizeof(int32_t)) != cudaSuccess) return set_error(MM_ERR_CUDA, "mmh: cudaMalloc failed");
        if (cudaHostAlloc((void**)&ws.h_counts, NSLOT * sizeof(int64_t), cudaHostAllocDefault) != cudaSuccess) return set_error(MM_ERR_CUDA, "mmh: cudaHostAlloc failed");
        ws.init = true;
    }
    if (bytes > ws.bytes) {
        if (ws.base) {                                    // nothing of ours is in flight here (every entry point drains its streams before it returns)
            for (int s = 0; s < NSLOT; ++s) cudaStreamSynchronize(ws.streams[s]);
            cudaFree(ws.base);
        }
        ws.base = nullptr; ws.bytes = 0;
        cudaError_t e = cudaMalloc(&ws.base, bytes);
        if (e != cudaSuccess) { cudaGetLastError(); return set_error(MM_ERR_CUDA, "mmh: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
        ws.bytes = bytes;
    }
    *out = &ws;
    return MM_OK;
}

// callers hold g_mu and have nothing in flight that touches the staging
int ensure_stage(Workspace& ws, size_t bytes) {
    if (bytes <= ws.h_stage_bytes) return MM_OK;
    if (ws.h_stage) cudaFreeHost(ws.h_stage);
    ws.h_stage = nullptr; ws.h_stage_bytes = 0;
    cudaError_t e = cudaHostAlloc((void**)&ws.h_stage, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return set_error(MM_ERR_CUDA, "mmh: cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
    ws.h_stage_bytes = bytes;
    return MM_OK;
}

// every exit path of an entry point — errors included — waits for the copies it enqueued: after the call returns nothing may still be
// reading the caller's input buffers or writing into its output buffers
int drain(Workspace& ws, int rc) {
    for (int s = 0; s < NSLOT; ++s) {
        cudaError_t e = cudaStreamSynchronize(ws.streams[s]);
        if (e != cudaSuccess && rc == MM_OK) { cudaGetLastError(); rc = set_error(MM_ERR_CUDA, "mmh: stream sync failed: %s", cudaGetErrorString(e)); }
    }
    return rc;
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
    if (e != cudaSuccess) { cudaGetLastError(); return set_error(MM_ERR_CUDA, "mmh: memcpy failed: %s", cudaGetErrorString(e)); }
    return MM_OK;
}

// a chunk of a device-resident array: AoS -> the slot pointer is the sub-range itself (nothing moves); SoA -> gather / scatter the nc
// component runs of the chunk into / out of the dense slot buffer
int resident_chunk(Arr& a, int slot, int64_t N, int64_t i0, int64_t n, int layout, bool load, cudaStream_t st) {
    char* res = (char*)a.d_res;
    if (layout != MM_LAYOUT_SOA || a.nc == 1) { a.d[slot] = res + (size_t)i0 * a.nc * a.elem; return MM_OK; }
    const size_t width = (size_t)n * a.elem, pitch = (size_t)N * a.elem;
    cudaError_t e = load ? cudaMemcpy2DAsync(a.d[slot], width, res + (size_t)i0 * a.elem, pitch, width, a.nc, cudaMemcpyDeviceToDevice, st)
                         : cudaMemcpy2DAsync(res + (size_t)i0 * a.elem, pitch, a.d[slot], width, width, a.nc, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) { cudaGetLastError(); return set_error(MM_ERR_CUDA, "mmh: resident-state staging failed: %s", cudaGetErrorString(e)); }
    return MM_OK;
}

// launch(slot, n, stream, d_fail) must enqueue the device-pointer entry point on the slot's chunk buffers
template <class Launch>
int pipeline(Arr* arrs, int narr, int64_t N, int layout, int32_t* n_fail, Launch launch) {
    if (N == 0) return MM_OK;
    std::lock_guard<std::mutex> lock(g_mu);              // host threads take turns, whichever mmh_* function they call
    const int64_t chunk = (N < g_chunk) ? N : g_chunk;
    size_t per_slot = 0;
    auto staged = [&](const Arr& a) { return a.h_in || a.h_out || (a.d_res && layout == MM_LAYOUT_SOA); };
    for (int a = 0; a < narr; ++a) if (staged(arrs[a])) per_slot += align256((size_t)chunk * arrs[a].nc * arrs[a].elem);
    Workspace* wsp = nullptr;
    int rc = ensure(per_slot * NSLOT, &wsp);
    if (rc) return rc;
    Workspace& ws = *wsp;
    char* base = (char*)ws.base;
    for (int s = 0; s < NSLOT; ++s) {
        size_t off = 0;
        for (int a = 0; a < narr; ++a) {
            if (staged(arrs[a])) { arrs[a].d[s] = base + s * per_slot + off; off += align256((size_t)chunk * arrs[a].nc * arrs[a].elem); }
            else arrs[a].d[s] = nullptr;
        }
    }
    if (cudaMemsetAsync(ws.d_fail, 0, sizeof(int32_t), ws.streams[0]) != cudaSuccess) return drain(ws, set_error(MM_ERR_CUDA, "mmh: memset failed"));
    if (cudaStreamSynchronize(ws.streams[0]) != cudaSuccess) return drain(ws, set_error(MM_ERR_CUDA, "mmh: sync failed"));
    int64_t k = 0;
    for (int64_t i0 = 0; i0 < N && rc == MM_OK; i0 += chunk, ++k) {
        const int slot = (int)(k % NSLOT);
        const int64_t n = (N - i0 < chunk) ? (N - i0) : chunk;
        cudaStream_t st = ws.streams[slot];
        for (int a = 0; a < narr && rc == MM_OK; ++a) {
            if (arrs[a].h_in) rc = copy_chunk(arrs[a], slot, N, i0, n, layout, true, st);
            else if (arrs[a].d_res) rc = resident_chunk(arrs[a], slot, N, i0, n, layout, true, st);
        }
        if (rc == MM_OK) rc = launch(slot, n, st, ws.d_fail);
        for (int a = 0; a < narr && rc == MM_OK; ++a) {
            if (arrs[a].h_out) rc = copy_chunk(arrs[a], slot, N, i0, n, layout, false, st);
            else if (arrs[a].d_res) rc = resident_chunk(arrs[a], slot, N, i0, n, layout, false, st);
        }
    }
    rc = drain(ws, rc);
    if (rc == MM_OK && n_fail) {
        int32_t f = 0;
        if (cudaMemcpy(&f, ws.d_fail, sizeof(f), cudaMemcpyDeviceToHost) != cudaSuccess) return set_error(MM_ERR_CUDA, "mmh: n_fail readback failed");
        *n_fail += f;
    }
    return rc;
}

inline Arr in_arr(const void* h, int nc, int elem = 8) { Arr a; memset(&a, 0, sizeof(a)); a.h_in = h; a.nc = nc; a.elem = elem; a.per_layout = true; return a; }
inline Arr res_arr(void* d, int nc, int elem = 8) { Arr a; memset(&a, 0, sizeof(a)); a.d_res = d; a.nc = nc; a.elem = elem; a.per_layout = true; return a; }
inline Arr out_arr(void* h, int nc, int elem = 8, bool per_layout = true) { Arr a; memset(&a, 0, sizeof(a)); a.h_out = h; a.nc = nc; a.elem = elem; a.per_layout = per_layout; return a; }

// ---- compact return of the J2 update (mm_compact.cu) behind the same chunk pipeline ----------------------------------------------
// Per chunk: H2D ε (+ state) -> mm_plastic_compact on the slot's buffers -> D2H row count (8 bytes, event) + σ / flags / iters (+ the dense
// state if asked for).  The rows are a variable-length copy, so they are issued LAG chunks later, when the host has seen the count:
// by then the next chunks' uploads and kernels are already queued, and neither copy engine waits for the host.
// `d_state` (resident mode): the state lives in device memory as one N-point array in `layout` and is updated in place — an SoA
// chunk is staged through the slot buffer with device-to-device 2-D copies, an AoS chunk is a contiguous sub-range used where it lies.
constexpr int LAG = 2;
static_assert(LAG < NSLOT, "a slot's row copy must be issued before the slot is reused");

// `expand_tan` (dense-through-factored mode, what mmh_plastic uses): the caller wants the DENSE tangent in host memory.  The factored rows
// {gξ, u, n} of every chunk (104 B per yielded point instead of 288 B per point over PCIe) land in an internal pinned staging together with
// the chunk's flags, and a helper thread expands each chunk into `expand_tan` with the host cores (mm_expand.cpp: the kernels' own fma
// expression, so the bits are those of the dense device path) while the next chunks are still in flight.  `rows` / `cap` are unused.
struct ExpandQueue {
    struct Job { int64_t j; int slot; int64_t c, i0, n; };
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::deque<Job> q;
    bool closed = false;
    int64_t done = 0;               // chunks [0, done) are expanded: their staging slots may be refilled
    int rc = MM_OK;
    char msg[256] = "";
};

int compact_pipeline(const MMPlastic* p, const double* eps, const double* state_in, double* d_state, double* sig, uint8_t* flags, uint16_t* iters,
                     int row_kind, double* rows, int64_t cap, int64_t* n_rows, double* state_out, int32_t* n_fail, const MMNewtonOpts* opts,
                     int64_t N, int layout, double* expand_tan = nullptr, int64_t chunk_cap = 0) {
    const int W = mm_plastic_row_width(row_kind);
    if (W < 0) return W;
    *n_rows = 0;
    if (N == 0) return MM_OK;
    std::lock_guard<std::mutex> lock(g_mu);
    int64_t chunk = (N < g_chunk) ? N : g_chunk;
    if (chunk_cap > 0 && chunk > chunk_cap) chunk = chunk_cap;
    const bool stage_state = !(d_state && layout == MM_LAYOUT_AOS);
    const size_t b_eps = align256((size_t)chunk * 48), b_st = stage_state ? align256((size_t)chunk * 112) : 0, b_sig = align256((size_t)chunk * 48),
                 b_fl = align256((size_t)chunk), b_it = align256((size_t)chunk * 2), b_rows = align256((size_t)chunk * W * 8), b_cnt = 256,
                 b_ws = align256((size_t)mm_plastic_compact_workspace(chunk, 1));
    const size_t per_slot = b_eps + b_st + b_sig + b_fl + b_it + b_rows + b_cnt + b_ws;
    Workspace* wsp = nullptr;
    int rc = ensure(per_slot * NSLOT, &wsp);
    if (rc) return rc;
    Workspace& ws = *wsp;
    struct Slot { char *eps, *st, *sig, *fl, *it, *rows, *cnt, *ws; } sl[NSLOT];
    for (int s = 0; s < NSLOT; ++s) {
        char* b = (char*)ws.base + s * per_slot;
        sl[s].eps = b; b += b_eps; sl[s].st = b; b += b_st; sl[s].sig = b; b += b_sig; sl[s].fl = b; b += b_fl; sl[s].it = b; b += b_it;
        sl[s].rows = b; b += b_rows; sl[s].cnt = b; b += b_cnt; sl[s].ws = b;
    }
    // dense-through-factored mode: pinned staging per slot = [rows: chunk x W doubles | flags: chunk bytes], expansion on a helper thread
    char* h_rows[NSLOT] = {};
    uint8_t* h_fl[NSLOT] = {};
    if (expand_tan) {
        const size_t s_rows = align256((size_t)chunk * W * 8), s_fl = align256((size_t)chunk);
        if ((rc = ensure_stage(ws, (s_rows + s_fl) * NSLOT)) != MM_OK) return rc;
        for (int s = 0; s < NSLOT; ++s) { h_rows[s] = ws.h_stage + s * (s_rows + s_fl); h_fl[s] = (uint8_t*)(h_rows[s] + s_rows); }
    }
    if (cudaMemsetAsync(ws.d_fail, 0, sizeof(int32_t), ws.streams[0]) != cudaSuccess) return drain(ws, set_error(MM_ERR_CUDA, "mmh: memset failed"));
    if (cudaStreamSynchronize(ws.streams[0]) != cudaSuccess) return drain(ws, set_error(MM_ERR_CUDA, "mmh: sync failed"));
    ExpandQueue xq;
    std::thread helper;
    if (expand_tan) {
        int dev = 0;
        cudaGetDevice(&dev);
        helper = std::thread([&, dev] {
            cudaSetDevice(dev);
            for (;;) {
                ExpandQueue::Job jb;
                {
                    std::unique_lock<std::mutex> lk(xq.mu);
                    xq.cv_job.wait(lk, [&] { return !xq.q.empty() || xq.closed; });
                    if (xq.q.empty()) return;
                    jb = xq.q.front();
                    xq.q.pop_front();
                }
                int r = MM_OK;
                const cudaError_t e = cudaEventSynchronize(ws.ev2[jb.slot]);      // rows, flags (and everything queued before them) are in host memory
                if (e != cudaSuccess) r = set_error(MM_ERR_CUDA, "mmh_plastic: waiting for a chunk's rows: %s", cudaGetErrorString(e));
                else r = expand_factored_chunk(p, h_fl[jb.slot], (const double*)h_rows[jb.slot], jb.c, expand_tan, N, jb.i0, jb.n, layout, 0);
                {
                    std::lock_guard<std::mutex> lk(xq.mu);
                    if (r != MM_OK && xq.rc == MM_OK) { xq.rc = r; strncpy(xq.msg, mm_last_error(), sizeof(xq.msg) - 1); }
                    xq.done = jb.j + 1;
                }
                xq.cv_done.notify_all();
            }
        });
    }
    // a staging slot is refilled by chunk j only after chunk j - NSLOT has been expanded out of it
    auto wait_expanded = [&](int64_t upto) -> int {
        if (!expand_tan) return MM_OK;
        std::unique_lock<std::mutex> lk(xq.mu);
        if (upto > 0) xq.cv_done.wait(lk, [&] { return xq.done >= upto || xq.rc != MM_OK; });
        return xq.rc;
    };
    auto close_helper = [&]() {
        if (!expand_tan || !helper.joinable()) return;
        { std::lock_guard<std::mutex> lk(xq.mu); xq.closed = true; }
        xq.cv_job.notify_all();
        helper.join();
    };
    Arr a_eps = in_arr(eps, 6), a_sin = in_arr(state_in, 14), a_sig = out_arr(sig, 6), a_fl = out_arr(flags, 1, 1, false), a_it = out_arr(iters, 1, 2, false),
        a_sout = out_arr(state_out, 14);
    for (int s = 0; s < NSLOT; ++s) { a_eps.d[s] = sl[s].eps; a_sin.d[s] = sl[s].st; a_sig.d[s] = sl[s].sig; a_fl.d[s] = sl[s].fl; a_it.d[s] = sl[s].it; a_sout.d[s] = sl[s].st; }
    const int64_t nchunks = (N + chunk - 1) / chunk;
    int64_t total = 0;
    auto fail = [&](const char* what, cudaError_t e) { cudaGetLastError(); return set_error(MM_ERR_CUDA, "mmh_plastic_compact: %s: %s", what, cudaGetErrorString(e)); };
    // the host learns chunk j's row count, then enqueues the copy of exactly that many rows behind everything already queued on the slot's stream
    auto finalize = [&](int64_t j) -> int {
        const int slot = (int)(j % NSLOT);
        cudaError_t e = cudaEventSynchronize(ws.ev[slot]);
        if (e != cudaSuccess) return fail("event sync", e);
        const int64_t c = ws.h_counts[slot];
        if (expand_tan) {
            if (c > 0 && (e = cudaMemcpyAsync(h_rows[slot], sl[slot].rows, (size_t)c * W * 8, cudaMemcpyDeviceToHost, ws.streams[slot])) != cudaSuccess) return fail("row copy", e);
            if ((e = cudaEventRecord(ws.ev2[slot], ws.streams[slot])) != cudaSuccess) return fail("event record", e);
            const int64_t i0 = j * chunk;
            { std::lock_guard<std::mutex> lk(xq.mu); xq.q.push_back({j, slot, c, i0, (N - i0 < chunk) ? (N - i0) : chunk}); }
            xq.cv_job.notify_one();
            total += c;
            return MM_OK;
        }
        if (c > 0 && total + c <= cap) {
            e = cudaMemcpyAsync(rows + total * W, sl[slot].rows, (size_t)c * W * 8, cudaMemcpyDeviceToHost, ws.streams[slot]);
            if (e != cudaSuccess) return fail("row copy", e);
        }
        total += c;                        // keeps counting past the capacity, so the caller learns the size it needs
        return MM_OK;
    };
    for (int64_t k = 0; k < nchunks && rc == MM_OK; ++k) {
        const int slot = (int)(k % NSLOT);
        const int64_t i0 = k * chunk, n = (N - i0 < chunk) ? (N - i0) : chunk;
        cudaStream_t st = ws.streams[slot];
        cudaError_t e = cudaSuccess;
        if (wait_expanded(k - NSLOT + 1) != MM_OK) break;        // the helper's error is picked up after the loop
        rc = copy_chunk(a_eps, slot, N, i0, n, layout, true, st);
        double* d_st = (double*)sl[slot].st;
        if (rc == MM_OK) {
            if (!d_state) rc = copy_chunk(a_sin, slot, N, i0, n, layout, true, st);
            else if (layout == MM_LAYOUT_AOS) d_st = d_state + i0 * 14;
            else if ((e = cudaMemcpy2DAsync(d_st, (size_t)n * 8, d_state + i0, (size_t)N * 8, (size_t)n * 8, 14, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) rc = fail("state staging", e);
        }
        if (rc == MM_OK)
            rc = mm_plastic_compact(p, (const double*)sl[slot].eps, d_st, (double*)sl[slot].sig, (uint8_t*)sl[slot].fl, iters ? (uint16_t*)sl[slot].it : nullptr, row_kind,
                                    (double*)sl[slot].rows, n, (int64_t*)sl[slot].cnt, d_st, ws.d_fail, opts, sl[slot].ws, b_ws, n, layout, st);
        if (rc == MM_OK && (e = cudaMemcpyAsync(&ws.h_counts[slot], sl[slot].cnt, sizeof(int64_t), cudaMemcpyDeviceToHost, st)) != cudaSuccess) rc = fail("count copy", e);
        if (rc == MM_OK && (e = cudaEventRecord(ws.ev[slot], st)) != cudaSuccess) rc = fail("event record", e);
        if (rc == MM_OK) rc = copy_chunk(a_sig, slot, N, i0, n, layout, false, st);
        if (rc == MM_OK && flags) rc = copy_chunk(a_fl, slot, N, i0, n, layout, false, st);
        if (rc == MM_OK && expand_tan && (e = cudaMemcpyAsync(h_fl[slot], sl[slot].fl, (size_t)n, cudaMemcpyDeviceToHost, st)) != cudaSuccess) rc = fail("flag staging", e);
        if (rc == MM_OK && iters) rc = copy_chunk(a_it, slot, N, i0, n, layout, false, st);
        if (rc == MM_OK && state_out) rc = copy_chunk(a_sout, slot, N, i0, n, layout, false, st);
        if (rc == MM_OK && d_state && layout == MM_LAYOUT_SOA &&
            (e = cudaMemcpy2DAsync(d_state + i0, (size_t)N * 8, d_st, (size_t)n * 8, (size_t)n * 8, 14, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) rc = fail("state write-back", e);
        if (rc == MM_OK && k >= LAG) rc = finalize(k - LAG);
    }
    for (int64_t j = (nchunks > LAG ? nchunks - LAG : 0); j < nchunks && rc == MM_OK; ++j) rc = finalize(j);
    close_helper();                        // the helper works its queue off (every queued job's event is recorded) and exits
    if (expand_tan && rc == MM_OK && xq.rc != MM_OK) rc = set_error(xq.rc, "%s", xq.msg);      // (joined: no lock needed)
    rc = drain(ws, rc);
    if (rc != MM_OK) return rc;
    *n_rows = total;
    if (expand_tan) cap = total;
    if (n_fail) {
        int32_t f = 0;
        if (cudaMemcpy(&f, ws.d_fail, sizeof(f), cudaMemcpyDeviceToHost) != cudaSuccess) return set_error(MM_ERR_CUDA, "mmh: n_fail readback failed");
        *n_fail += f;
    }
    if (total > cap) return set_error(MM_ERR_CAPACITY, "mmh_plastic_compact: %lld points yielded but `rows` holds %lld rows; sigma / flags / iters are complete, call again with a larger buffer",
                                      (long long)total, (long long)cap);
    return MM_OK;
}
}  // namespace

// run-time switches of the host-buffer entry points (mm_set_tuning): "mmh_plastic_via_factored" 1 (default) / 0
int* host_tuning_slot(const char* name) {
    static int via_factored = [] { const char* e = getenv("MM_MMH_PLASTIC_VIA_FACTORED"); return (e && e[0] == '0') ? 0 : 1; }();
    if (!strcmp(name, "mmh_plastic_via_factored")) return &via_factored;
    return nullptr;
}
}  // namespace mm

using namespace mm;

extern "C" {

void* mmh_alloc_pinned(uint64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); set_error(MM_ERR_CUDA, "mmh_alloc_pinned(%llu) failed", (unsigned long long)bytes); return nullptr; }
    return p;
}
void mmh_free_pinned(void* p) { if (p) cudaFreeHost(p); }
int64_t mmh_set_chunk_points(int64_t n) { std::lock_guard<std::mutex> lock(g_mu); int64_t old = g_chunk; if (n > 0) g_chunk = n; return old; }

int mmh_linear_elastic(const MMLinearElastic* p, const double* eps, double* sig, double* tan_or_null, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!eps || !sig))) return set_error(MM_ERR_ARG, "mmh_linear_elastic: bad argument");
    // the tangent is the one stored Eᵉ for every point (src/LinearElastic.jl:59): written by the host cores while ε and σ cross the link,
    // instead of 288 bytes per point coming down ("mmh_plastic_via_factored" 0 also restores the plain transfer here); same values — the
    // kernel stores iso_entry(λ, G) of the same host-computed λ, G
    const bool host_fill = tan_or_null && N >= (1 << 14) && *host_tuning_slot("mmh_plastic_via_factored") != 0 && (layout == MM_LAYOUT_SOA || layout == MM_LAYOUT_AOS);
    std::thread filler;
    if (host_fill) {
        const double nu = p->nu, E = p->E;
        const double lam = E * nu / ((1 + nu) * (1 - 2 * nu)), G = E / (2 * (1 + nu));          // make_elastic_k (mm_math.cuh)
        filler = std::thread([=] { fill_constant_tangent(lam, G, tan_or_null, N, layout, 0); });
    }
    Arr a[3] = {in_arr(eps, 6), out_arr(sig, 6), out_arr(host_fill ? nullptr : tan_or_null, 36)};
    const int rc = pipeline(a, 3, N, layout, nullptr, [&](int s, int64_t n, cudaStream_t st, int32_t* df) {
        return mm_linear_elastic(p, (const double*)a[0].d[s], (double*)a[1].d[s], (double*)a[2].d[s], n, layout, st);
    });
    if (filler.joinable()) filler.join();
    return rc;
}

int mmh_plastic(const MMPlastic* p, const double* eps, const double* state_in, double* sig, double* tan, double* state_out,
                uint8_t* flags, uint16_t* iters, int32_t* n_fail, const MMNewtonOpts* opts, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!eps || !state_in || !sig || !state_out))) return set_error(MM_ERR_ARG, "mmh_plastic: bad argument");
    // dense tangent wanted: send the factored rows of the yielded points over the link and let the host cores expand them chunk by chunk
    // while the next chunks are in flight — same bits as the dense device path (tests/test_compact.py), 217 instead of 451 bytes per
    // point coming down ("mmh_plastic_via_factored" 0 restores the plain dense transfer)
    // The expansion only pays when it overlaps the transfer of later chunks: batches below 2^18 points go the plain way (measured: 111 vs 228 us
    // at 10^3 points, 1.25 vs 2.15 ms at 10^5), medium batches are cut into >= 8 chunks (profiles/r02_host_path_dense.log).  "...via_factored" 2
    // forces the mode for any size (tests).
    const int via = *host_tuning_slot("mmh_plastic_via_factored");
    if (tan && N > 0 && via != 0 && (N >= (1 << 18) || via == 2) && (layout == MM_LAYOUT_SOA || layout == MM_LAYOUT_AOS)) {
        int64_t n_rows = 0, cap8 = N / 8;
        if (cap8 < (1 << 16)) cap8 = 1 << 16;
        return compact_pipeline(p, eps, state_in, nullptr, sig, flags, iters, MM_ROWS_FACTORED, nullptr, 0, &n_rows, state_out, n_fail, opts, N, layout, tan, cap8);
    }
    Arr a[7] = {in_arr(eps, 6), in_arr(state_in, 14), out_arr(sig, 6), out_arr(tan, 36), out_arr(state_out, 14),
                out_arr(flags, 1, 1, false), out_arr(iters, 1, 2, false)};
    return pipeline(a, 7, N, layout, n_fail, [&](int s, int64_t n, cudaStream_t st, int32_t* df) {
        return mm_plastic(p, (const double*)a[0].d[s], (const double*)a[1].d[s], (double*)a[2].d[s], (double*)a[3].d[s], (double*)a[4].d[s],
                          (uint8_t*)a[5].d[s], (uint16_t*)a[6].d[s], df, opts, n, layout, st);
    });
}

int mmh_hyper(int model, const MMHyper* p, const double* strain, int strain_kind, double* S, double* dSdC, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!strain || !S))) return set_error(MM_ERR_ARG, "mmh_hyper: bad argument");
    Arr a[3] = {in_arr(strain, strain_kind == MM_STRAIN_F ? 9 : 6), out_arr(S, 6), out_arr(dSdC, 36)};
    return pipeline(a, 3, N, layout, nullptr, [&](int s, int64_t n, cudaStream_t st, int32_t* df) {
        return mm_hyper(model, p, (const double*)a[0].d[s], strain_kind, (double*)a[1].d[s], (double*)a[2].d[s], n, layout, st);
    });
}

int mmh_crystal(const MMCrystal* p, const double* deps, double dt, const double* state_in, double* sig, double* tan, double* state_out,
                uint16_t* active, uint16_t* iters, int32_t* n_fail, const MMNewtonOpts* opts, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!deps || !state_in || !sig || !state_out))) return set_error(MM_ERR_ARG, "mmh_crystal: bad argument");
    Arr a[7] = {in_arr(deps, 6), in_arr(state_in, 42), out_arr(sig, 6), out_arr(tan, 36), out_arr(state_out, 42),
                out_arr(active, 1, 2, false), out_arr(iters, 1, 2, false)};
    return pipeline(a, 7, N, layout, n_fail, [&](int s, int64_t n, cudaStream_t st, int32_t* df) {
        return mm_crystal(p, (const double*)a[0].d[s], dt, (const double*)a[1].d[s], (double*)a[2].d[s], (double*)a[3].d[s], (double*)a[4].d[s],
                          (uint16_t*)a[5].d[s], (uint16_t*)a[6].d[s], df, opts, n, layout, st);
    });
}

int mmh_crystal_resident(const MMCrystal* p, const double* deps, double dt, double* d_state, double* sig, double* tan, uint16_t* active, uint16_t* iters,
                         int32_t* n_fail, const MMNewtonOpts* opts, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!deps || !d_state || !sig))) return set_error(MM_ERR_ARG, "mmh_crystal_resident: bad argument");
    if (layout != MM_LAYOUT_SOA && layout != MM_LAYOUT_AOS) return set_error(MM_ERR_ARG, "mmh_crystal_resident: unknown layout %d", layout);
    Arr a[6] = {in_arr(deps, 6), res_arr(d_state, 42), out_arr(sig, 6), out_arr(tan, 36), out_arr(active, 1, 2, false), out_arr(iters, 1, 2, false)};
    return pipeline(a, 6, N, layout, n_fail, [&](int s, int64_t n, cudaStream_t st, int32_t* df) {
        return mm_crystal(p, (const double*)a[0].d[s], dt, (const double*)a[1].d[s], (double*)a[2].d[s], (double*)a[3].d[s], (double*)a[1].d[s],
                          (uint16_t*)a[4].d[s], (uint16_t*)a[5].d[s], df, opts, n, layout, st);
    });
}

int mmh_xu_needleman(const MMXuNeedleman* p, const double* jump, double* T, double* dTdD, int dim, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!jump || !T)) || (dim != 2 && dim != 3)) return set_error(MM_ERR_ARG, "mmh_xu_needleman: bad argument");
    Arr a[3] = {in_arr(jump, dim), out_arr(T, dim), out_arr(dTdD, dim * dim)};
    return pipeline(a, 3, N, layout, nullptr, [&](int s, int64_t n, cudaStream_t st, int32_t* df) {
        return mm_xu_needleman(p, (const double*)a[0].d[s], (double*)a[1].d[s], (double*)a[2].d[s], dim, n, layout, st);
    });
}

int mmh_wrapped_linear_elastic(int dim_kind, const MMLinearElastic* p, const double* eps_red, double* sig_red, double* tan_red, uint8_t* outer_iters,
                               int32_t* n_fail, const MMNewtonOpts* wrapper_opts, int64_t N, int layout) {
    const int nc = mm_wrap_ncomp(dim_kind);
    if (nc < 0) return nc;
    if (!p || N < 0 || (N > 0 && (!eps_red || !sig_red))) return set_error(MM_ERR_ARG, "mmh_wrapped_linear_elastic: bad argument");
    Arr a[4] = {in_arr(eps_red, nc), out_arr(sig_red, nc), out_arr(tan_red, nc * nc), out_arr(outer_iters, 1, 1, false)};
    return pipeline(a, 4, N, layout, n_fail, [&](int s, int64_t n, cudaStream_t st, int32_t* df) {
        return mm_wrapped_linear_elastic(dim_kind, p, (const double*)a[0].d[s], (double*)a[1].d[s], (double*)a[2].d[s], (uint8_t*)a[3].d[s], df,
                                         wrapper_opts, n, layout, st);
    });
}

int mmh_wrapped_plastic(int dim_kind, const MMPlastic* p, const double* eps_red, const double* state_in, double* sig_red, double* tan_red, double* state_out,
                        uint8_t* flags, uint16_t* iters, uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* opts, const MMNewtonOpts* wrapper_opts,
                        int64_t N, int layout) {
    const int nc = mm_wrap_ncomp(dim_kind);
    if (nc < 0) return nc;
    if (!p || N < 0 || (N > 0 && (!eps_red || !state_in || !sig_red || !state_out))) return set_error(MM_ERR_ARG, "mmh_wrapped_plastic: bad argument");
    Arr a[8] = {in_arr(eps_red, nc), in_arr(state_in, 14), out_arr(sig_red, nc), out_arr(tan_red, nc * nc), out_arr(state_out, 14),
                out_arr(flags, 1, 1, false), out_arr(iters, 1, 2, false), out_arr(outer_iters, 1, 1, false)};
    return pipeline(a, 8, N, layout, n_fail, [&](int s, int64_t n, cudaStream_t st, int32_t* df) {
        return mm_wrapped_plastic(dim_kind, p, (const double*)a[0].d[s], (const double*)a[1].d[s], (double*)a[2].d[s], (double*)a[3].d[s], (double*)a[4].d[s],
                                  (uint8_t*)a[5].d[s], (uint16_t*)a[6].d[s], (uint8_t*)a[7].d[s], df, opts, wrapper_opts, n, layout, st);
    });
}

int mmh_wrapped_crystal(int dim_kind, const MMCrystal* p, const double* deps_red, double dt, const double* state_in, double* sig_red, double* tan_red,
                        double* state_out, uint16_t* active, uint16_t* iters, uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* opts,
                        const MMNewtonOpts* wrapper_opts, int64_t N, int layout) {
    const int nc = mm_wrap_ncomp(dim_kind);
    if (nc < 0) return nc;
    if (!p || N < 0 || (N > 0 && (!deps_red || !state_in || !sig_red || !state_out))) return set_error(MM_ERR_ARG, "mmh_wrapped_crystal: bad argument");
    Arr a[8] = {in_arr(deps_red, nc), in_arr(state_in, 42), out_arr(sig_red, nc), out_arr(tan_red, nc * nc), out_arr(state_out, 42),
                out_arr(active, 1, 2, false), out_arr(iters, 1, 2, false), out_arr(outer_iters, 1, 1, false)};
    return pipeline(a, 8, N, layout, n_fail, [&](int s, int64_t n, cudaStream_t st, int32_t* df) {
        return mm_wrapped_crystal(dim_kind, p, (const double*)a[0].d[s], dt, (const double*)a[1].d[s], (double*)a[2].d[s], (double*)a[3].d[s],
                                  (double*)a[4].d[s], (uint16_t*)a[5].d[s], (uint16_t*)a[6].d[s], (uint8_t*)a[7].d[s], df, opts, wrapper_opts, n,
                                  layout, st);
    });
}

int mmh_wrapped_transversely_isotropic(int dim_kind, const MMTransverselyIsotropic* p, const double* eps_red, const double* a3, double* sig_red,
                                       double* tan_red, uint8_t* outer_iters, int32_t* n_fail, const MMNewtonOpts* wrapper_opts, int64_t N, int layout) {
    const int nc = mm_wrap_ncomp(dim_kind);
    if (nc < 0) return nc;
    if (!p || N < 0 || (N > 0 && (!eps_red || !a3 || !sig_red))) return set_error(MM_ERR_ARG, "mmh_wrapped_transversely_isotropic: bad argument");
    Arr a[5] = {in_arr(eps_red, nc), in_arr(a3, 3), out_arr(sig_red, nc), out_arr(tan_red, nc * nc), out_arr(outer_iters, 1, 1, false)};
    return pipeline(a, 5, N, layout, n_fail, [&](int s, int64_t n, cudaStream_t st, int32_t* df) {
        return mm_wrapped_transversely_isotropic(dim_kind, p, (const double*)a[0].d[s], (const double*)a[1].d[s], (double*)a[2].d[s], (double*)a[3].d[s],
                                                 (uint8_t*)a[4].d[s], df, wrapper_opts, n, layout, st);
    });
}

int mmh_hyper_ex(int model, const MMHyper* p, const double* strain, int strain_kind, int tangent_kind, double* stress, double* tangent, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!strain || !stress))) return set_error(MM_ERR_ARG, "mmh_hyper_ex: bad argument");
    const bool pf = tangent_kind == MM_TANGENT_DPDF;
    Arr a[3] = {in_arr(strain, strain_kind == MM_STRAIN_F ? 9 : 6), out_arr(stress, pf ? 9 : 6), out_arr(tangent, pf ? 81 : 36)};
    return pipeline(a, 3, N, layout, nullptr, [&](int s, int64_t n, cudaStream_t st, int32_t* df) {
        return mm_hyper_ex(model, p, (const double*)a[0].d[s], strain_kind, tangent_kind, (double*)a[1].d[s], (double*)a[2].d[s], n, layout, st);
    });
}

int mmh_transversely_isotropic(const MMTransverselyIsotropic* p, const double* eps, const double* a3, double* sig, double* tan, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!eps || !a3 || !sig))) return set_error(MM_ERR_ARG, "mmh_transversely_isotropic: bad argument");
    Arr a[4] = {in_arr(eps, 6), in_arr(a3, 3), out_arr(sig, 6), out_arr(tan, 36)};
    return pipeline(a, 4, N, layout, nullptr, [&](int s, int64_t n, cudaStream_t st, int32_t* df) {
        return mm_transversely_isotropic(p, (const double*)a[0].d[s], (const double*)a[1].d[s], (double*)a[2].d[s], (double*)a[3].d[s], n, layout, st);
    });
}

int mmh_linear_elastic_energy(const MMLinearElastic* p, const double* eps, double* psi, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!eps || !psi))) return set_error(MM_ERR_ARG, "mmh_linear_elastic_energy: bad argument");
    Arr a[2] = {in_arr(eps, 6), out_arr(psi, 1, 8, false)};
    return pipeline(a, 2, N, layout, nullptr, [&](int s, int64_t n, cudaStream_t st, int32_t* df) {
        return mm_linear_elastic_energy(p, (const double*)a[0].d[s], (double*)a[1].d[s], n, layout, st);
    });
}

int mmh_hyper_energy(int model, const MMHyper* p, const double* strain, int strain_kind, double* psi, int64_t N, int layout) {
    if (!p || N < 0 || (N > 0 && (!strain || !psi))) return set_error(MM_ERR_ARG, "mmh_hyper_energy: bad argument");
    Arr a[2] = {in_arr(strain, strain_kind == MM_STRAIN_F ? 9 : 6), out_arr(psi, 1, 8, false)};
    return pipeline(a, 2, N, layout, nullptr, [&](int s, int64_t n, cudaStream_t st, int32_t* df) {
        return mm_hyper_energy(model, p, (const double*)a[0].d[s], strain_kind, (double*)a[1].d[s], n, layout, st);
    });
}

int mmh_plastic_compact(const MMPlastic* p, const double* eps, const double* state_in, double* sig, uint8_t* flags, uint16_t* iters, int row_kind,
                        double* rows, int64_t rows_capacity, int64_t* n_rows, double* state_out, int32_t* n_fail, const MMNewtonOpts* opts, int64_t N, int layout) {
    if (!p || N < 0 || rows_capacity < 0 || !n_rows || (N > 0 && (!eps || !state_in || !sig || !flags || (rows_capacity > 0 && !rows))))
        return set_error(MM_ERR_ARG, "mmh_plastic_compact: bad argument");
    if (layout != MM_LAYOUT_SOA && layout != MM_LAYOUT_AOS) return set_error(MM_ERR_ARG, "mmh_plastic_compact: unknown layout %d", layout);
    return compact_pipeline(p, eps, state_in, nullptr, sig, flags, iters, row_kind, rows, rows_capacity, n_rows, state_out, n_fail, opts, N, layout);
}

int mmh_plastic_compact_resident(const MMPlastic* p, const double* eps, double* d_state, double* sig, uint8_t* flags, uint16_t* iters, int row_kind,
                                 double* rows, int64_t rows_capacity, int64_t* n_rows, int32_t* n_fail, const MMNewtonOpts* opts, int64_t N, int layout) {
    if (!p || N < 0 || rows_capacity < 0 || !n_rows || (N > 0 && (!eps || !d_state || !sig || !flags || (rows_capacity > 0 && !rows))))
        return set_error(MM_ERR_ARG, "mmh_plastic_compact_resident: bad argument");
    if (layout != MM_LAYOUT_SOA && layout != MM_LAYOUT_AOS) return set_error(MM_ERR_ARG, "mmh_plastic_compact_resident: unknown layout %d", layout);
    if (row_kind != MM_ROWS_TANGENT && row_kind != MM_ROWS_FACTORED)
        return set_error(MM_ERR_ARG, "mmh_plastic_compact_resident: the state stays on the device; row kind must be MM_ROWS_TANGENT or MM_ROWS_FACTORED");
    return compact_pipeline(p, eps, nullptr, d_state, sig, flags, iters, row_kind, rows, rows_capacity, n_rows, nullptr, n_fail, opts, N, layout);
}

}  // extern "C"
