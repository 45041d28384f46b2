// wfm_api.cu -- implementation of the C ABI in include/wfm_b200.h.
//
// Host-side state machine of the reference class (PState / freeMem protocol, setter order,
// IllegalArgumentException sites) plus the kernel launches.  WFM = WideFieldModel.java.
#include "../../include/wfm_b200.h"
#include "wfm_kernels.cuh"
#include "wfm_generic.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>
#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <vector>

using namespace wfm;

namespace {
thread_local std::string g_create_error;
}  // namespace

namespace wfm_detail {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need) {
        if (need <= bytes && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMalloc(&p, need ? need : 1);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

}  // namespace wfm_detail
using wfm_detail::DevBuf;

struct wfm_model {
    // geometry (MicroscopeModel.java:62-78, WFM:154-172)
    int N = 0, nz_global = 0, z0 = 0, nzl = 0;
    // batch of independent models (wfm_create_batch): nbatch models of nzm planes each, stacked: nzl = nbatch*nzm.
    // They share optics, basis and support; rho/phi/psi/mask and the coefficient vectors are per model.
    int nbatch = 1, nzm = 0;
    std::vector<double> alpha_b, beta_b, bpar_h;  // [nbatch][nphase], [nbatch][nmod], [nbatch][4] = {ni/lambda, dX, dY, 1/|beta|}
    double dxy = 0, dz = 0;
    int precision = WFM_F64;
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // second stream for wfm_get_psf_async: the D2H of the PSF runs beside the H2D of q (full-duplex PCIe)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_copied = nullptr, ev_order = nullptr;
    bool copy_pending = false;
    // host-buffer entry points (wfm_get_psf_async, wfm_apply_j_* with host q): the slab moves in plane chunks so that
    // the PCIe copies overlap the kernels -- a launch then covers the plane window [win0, win0 + winN) only
    // (winN == 0: the whole slab); in_stream carries the H2D chunks of q, ev_chunk[c] marks chunk c
    int win0 = 0, winN = 0;
    cudaStream_t in_stream = nullptr;
    // pageable host arrays (Java-heap double[] behind GetPrimitiveArrayCritical, numpy): multi-threaded staging (HostStage)
    struct HostStage* stage = nullptr;
    int siblings = 1;                             // children of a multi-device handle share the host's cores
    std::vector<cudaEvent_t> ev_chunk, ev_in;
    // optics (WFM:161-166)
    bool have_optics = false;
    double NA = 0, lambda = 0, ni = 0, lambda_ni = 0, radius = 0, deltaX = 0, deltaY = 0;
    int ndefocus = 3;
    // basis
    int nzern = 0, radial = 0;
    DevBuf Z;
    // coefficients (parameterCoefs[], MicroscopeModel.java:54)
    int nphase = 0, nmod = 0;
    Coefs alpha{}, beta{};
    bool have_rho = false;
    // pupil arrays
    DevBuf rho, phi, psi, mask, map, support;
    std::vector<uint8_t> h_map, h_zsup, h_esc;
    bool activity_dirty = true;
    int nax = 0, nay = 0, pitch = 0, ctile = 1;
    bool narrow = false;                          // support inside [0,N/4) u [3N/4,N) on both axes ("narrow" kernels)
    DevBuf act_x, inv_x, act_y, inv_y, cell_list, in_list, Zs;
    bool basis_packed = false;
    DevBuf s_rho, s_phi, s_psi, s_flags;          // pupil strip [N][pitch]
    bool strip_dirty = true;
    bool phi_clean_off_support = false;           // phi == 0 off the mask (left so by a full k_set_phase; the escape hatch clears it)
    int ncells = 0;
    // FFT twiddles
    DevBuf tw;
    DevBuf cis_tab;                               // (cos, sin)(2 pi k/64), k < 64: table of wfm_cis
    bool generic = false;                         // N is not one of the pipeline plans: any-N path (wfm_generic.cuh)
    DevBuf tw64;                                  // any-N path: W_N^m in double (also in fp32 mode)
    // outputs + PState (MicroscopeModel.java:42)
    DevBuf cpx, psf;
    int pstate = 0;
    // scratch
    DevBuf scratch, Gj, Gm, ctl, block_part, grad, qdev;
    DevBuf alpha_dev, beta_dev, bpar_dev;         // device copies of the batch tables
    int num_sms = 148;
    unsigned long pipe_checks = 0;
    bool ctl_dirty = true;
    int modulus_mode = WFM_MODULUS_INTENDED;
    std::string err;
    // optional per-kernel CUDA-event timing (wfm_set_profiling)
    bool profiling = false;
    struct Span { int kid; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> free_events;
    double k_ms[WFM_KERNEL_IDS] = {0};
    uint64_t k_count[WFM_KERNEL_IDS] = {0};

    // multi-device handle (wfm_create_multi): one z-slab child per device, this object holds no device memory of its
    // own except the landing buffer of the partial gradient vectors on the first device
    std::vector<wfm_model*> parts;
    std::vector<int> part_z0;                     // first plane of each child inside the stack
    DevBuf xslots;                                // [n_dev][glen] doubles on parts[0]'s device, written by the peers
    std::vector<int> peer_direct;                 // child i can store into xslots over NVLink (peer access enabled)
    std::vector<cudaEvent_t> part_done;           // one event per child: its Jacobian chain has been queued
    // cross-process gradient exchange over CUDA IPC peer memory (wfm_exchange_*): see wfm_multi.inl
    struct Exchange* xchg = nullptr;
    bool multi() const { return !parts.empty(); }

    int npix() const { return N * N; }
    size_t esz() const { return precision == WFM_F64 ? 8 : 4; }
    int glen() const { return 3 + nphase + nmod; }
    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
        err = buf;
        return code;
    }
};

#define WFM_CK(h, call)                                                                         \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return (h)->fail(e__ == cudaErrorMemoryAllocation ? WFM_ERR_NOMEM : WFM_ERR_CUDA,   \
                             "%s failed: %s", #call, cudaGetErrorString(e__));                  \
    } while (0)

#define WFM_CK_LAUNCH(h, what)                                                                  \
    do {                                                                                        \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess)                                                                 \
            return (h)->fail(WFM_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e__)); \
    } while (0)

// Every entry point that allocates, copies or launches runs with the handle's device current and hands the
// caller's device back on return: a fresh host thread starts on device 0, whatever device its handle lives on.
struct DeviceScope {
    int prev = -1;
    bool switched = false, ok = true;
    explicit DeviceScope(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) { ok = cudaSetDevice(dev) == cudaSuccess; switched = ok && prev >= 0; }
    }
    ~DeviceScope() { if (switched) cudaSetDevice(prev); }
    DeviceScope(const DeviceScope&) = delete;
    DeviceScope& operator=(const DeviceScope&) = delete;
};
#define WFM_ENTER(h)                                                                             \
    DeviceScope dev_scope__((h)->device);                                                        \
    if (!dev_scope__.ok) return (h)->fail(WFM_ERR_CUDA, "cudaSetDevice(%d) failed", (h)->device)

// Cross-process gradient exchange state (one process per GPU; wfm_exchange_export / _connect, wfm_multi.inl).
struct Exchange {
    int rank = -1, world = 0, glen_cap = 0;
    void* base = nullptr;                    // this rank's landing buffer (cudaMalloc): slots, then flags
    size_t slot_bytes = 0;
    std::vector<void*> mapped;               // every rank's buffer as mapped into this process (own = base)
    DevBuf local;                            // ticket + err words
    unsigned epoch = 0;
    bool connected = false;
};

// ---- pageable host arrays: staged, multi-threaded copies ---------------------------------------------------------
// cudaMemcpy from / to ordinary (pageable) memory goes through the driver's bounce buffer on ONE thread: ~6-7 GB/s here
// (78 ms for the 2 x 537 MB of a 512^2 x 256 step against 12 ms from pinned buffers).  TiPi's arrays are Java-heap
// double[]: whatever the binding does, the library sees pageable memory.  So for large pageable arrays the
// library stages itself: T host threads copy their share of every PIECE between the caller's array and pinned slots
// (two per thread: one being filled or drained by the thread, one on the PCIe link), and the device copies of the
// pieces are queued as they become ready.
struct HostStage {
    int threads = 0;
    size_t share = 0;                          // bytes per thread and slot
    std::vector<void*> slot;                   // [threads][2], pinned
    std::vector<cudaEvent_t> ev;               // [threads][2]: the slot's last device copy has completed
    size_t piece() const { return share * (size_t)threads; }
    void release() {
        for (void* p : slot) if (p) cudaFreeHost(p);
        for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
        slot.clear(); ev.clear(); threads = 0;
    }
};

namespace {

bool host_is_pageable(const void* p) {
    if (getenv("WFM_FORCE_STAGED")) return true;           // (tests: take the staged path for any pointer)
    if (getenv("WFM_NO_STAGED")) return false;
#ifdef WFM_EMU
    (void)p;
    return false;
#else
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
#endif
}

// the staging area of a handle (created on first use, kept): NULL when staging is off or cannot be set up
HostStage* host_stage(wfm_model* h) {
    if (h->stage) return h->stage->threads > 0 ? h->stage : nullptr;
    h->stage = new HostStage();
    int t = (int)std::thread::hardware_concurrency() / (2 * (h->siblings > 0 ? h->siblings : 1));
    if (t > 8) t = 8;
    if (const char* e = getenv("WFM_HOST_THREADS")) t = atoi(e);
    if (t < 1) t = 1;
    size_t share = (size_t)1 << 20;               // (measured, 2 x 537 MB per step: 256 KB 48.9 ms, 512 KB 41.3, 1 MB 27.0, 2 MB 27.9, 4 MB 28.9, 16 MB 33.7)
    if (const char* e = getenv("WFM_HOST_SHARE_BYTES")) share = (size_t)atoll(e);          // (tests: many small pieces)
    share = (share + 255) / 256 * 256;
    HostStage* st = h->stage;
    st->slot.assign((size_t)2 * t, nullptr); st->ev.assign((size_t)2 * t, nullptr);
    for (int i = 0; i < 2 * t; ++i)
        if (cudaHostAlloc(&st->slot[i], share, cudaHostAllocDefault) != cudaSuccess ||
            cudaEventCreateWithFlags(&st->ev[i], cudaEventDisableTiming) != cudaSuccess) { st->release(); cudaGetLastError(); return nullptr; }
    st->threads = t; st->share = share;
    return st;
}

// (no exception may cross the C ABI: a thread that cannot be started is reported, its share is done by the caller)
template <class W> bool spawn_worker(std::vector<std::thread>& pool, W& work, int w) {
    try { pool.emplace_back([&work, w] { work(w); }); return true; } catch (...) { return false; }
}

// Host -> device through the staging slots, on `stream`.  Returns after every piece has been QUEUED; on_piece(k) is
// called on the calling thread, in order, once piece k (bytes [k*piece, (k+1)*piece)) is completely queued.
template <class F> int staged_h2d(wfm_model* h, HostStage* st, void* dev, const void* host, size_t bytes, cudaStream_t stream,
                                  F on_piece) {
    const int T = st->threads;
    const size_t piece = st->piece();
    const int npieces = (int)((bytes + piece - 1) / piece);
    std::unique_ptr<std::atomic<int>[]> arrived(new std::atomic<int>[npieces]);
    for (int k = 0; k < npieces; ++k) arrived[k].store(0);
    std::atomic<int> failed{0};
    const int device = h->device;
    auto work = [&](int w) {
        if (cudaSetDevice(device) != cudaSuccess) failed.store(1);
        for (int k = 0; k < npieces; ++k) {
            const int s = 2 * w + (k & 1);
            const size_t off = (size_t)k * piece + (size_t)w * st->share;
            if (off < bytes && !failed.load()) {
                const size_t len = std::min(st->share, bytes - off);
                if (k >= 2 && cudaEventSynchronize(st->ev[s]) != cudaSuccess) failed.store(1);   // the slot's previous copy has left
                memcpy(st->slot[s], (const char*)host + off, len);
                if (cudaMemcpyAsync((char*)dev + off, st->slot[s], len, cudaMemcpyHostToDevice, stream) != cudaSuccess ||
                    cudaEventRecord(st->ev[s], stream) != cudaSuccess) failed.store(1);
            }
            arrived[k].fetch_add(1);
        }
    };
    std::vector<std::thread> pool;
    std::vector<int> inline_w;                       // shares whose thread could not be started: done by the caller itself
    for (int w = 0; w < T; ++w)
        if (T == 1 || !spawn_worker(pool, work, w)) inline_w.push_back(w);
    for (int w : inline_w) work(w);
    int rc = WFM_OK;
    for (int k = 0; k < npieces; ++k) {
        while (arrived[k].load() < T) std::this_thread::yield();
        if (!rc) rc = on_piece(k);
    }
    for (auto& t : pool) t.join();
    if (failed.load()) { cudaGetLastError(); return h->fail(WFM_ERR_CUDA, "staged host -> device copy failed"); }
    return rc;
}

// Device -> host through the staging slots, on `stream`; blocks until `host` holds all the bytes.  ready(b) must return
// (after queueing what `stream` has to wait for) once the device bytes [0, b) may be read; it is called by the worker
// threads with increasing b.
template <class F> int staged_d2h(wfm_model* h, HostStage* st, void* host, const void* dev, size_t bytes, cudaStream_t stream,
                                  F ready) {
    const int T = st->threads;
    const size_t piece = st->piece();
    const int npieces = (int)((bytes + piece - 1) / piece);
    std::atomic<int> failed{0};
    const int device = h->device;
    auto work = [&](int w) {
        if (cudaSetDevice(device) != cudaSuccess) failed.store(1);
        auto queue = [&](int k) {
            const size_t off = (size_t)k * piece + (size_t)w * st->share;
            if (off >= bytes || failed.load()) return;
            const size_t len = std::min(st->share, bytes - off);
            const int s = 2 * w + (k & 1);
            if (ready(off + len) != WFM_OK) { failed.store(1); return; }
            if (cudaMemcpyAsync(st->slot[s], (const char*)dev + off, len, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
                cudaEventRecord(st->ev[s], stream) != cudaSuccess) failed.store(1);
        };
        queue(0);
        for (int k = 0; k < npieces; ++k) {
            if (k + 1 < npieces) queue(k + 1);                   // the other slot: on the link while this one is drained
            const size_t off = (size_t)k * piece + (size_t)w * st->share;
            if (off >= bytes || failed.load()) continue;
            const size_t len = std::min(st->share, bytes - off);
            const int s = 2 * w + (k & 1);
            if (cudaEventSynchronize(st->ev[s]) != cudaSuccess) { failed.store(1); continue; }
            memcpy((char*)host + off, st->slot[s], len);
        }
    };
    std::vector<std::thread> pool;
    std::vector<int> inline_w{0};
    for (int w = 1; w < T; ++w)
        if (!spawn_worker(pool, work, w)) inline_w.push_back(w);
    for (int w : inline_w) work(w);
    for (auto& t : pool) t.join();
    if (failed.load()) { cudaGetLastError(); return h->fail(WFM_ERR_CUDA, "staged device -> host copy failed"); }
    return WFM_OK;
}

cudaEvent_t take_event(wfm_model* h) {
    if (!h->free_events.empty()) { cudaEvent_t e = h->free_events.back(); h->free_events.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
// Brackets the launches issued in its scope with two events on the handle's stream.
struct KernelSpan {
    wfm_model* h; int kid; cudaEvent_t a = nullptr;
    KernelSpan(wfm_model* h_, int kid_) : h(h_), kid(kid_) {
        if (h->profiling) { a = take_event(h); cudaEventRecord(a, h->stream); }
    }
    ~KernelSpan() {
        if (a) { cudaEvent_t b = take_event(h); cudaEventRecord(b, h->stream); h->spans.push_back({kid, a, b}); }
    }
};
void drain_spans(wfm_model* h) {
    for (auto& sp : h->spans) {
        float ms = 0.f;
        cudaEventSynchronize(sp.b);
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) { h->k_ms[sp.kid] += ms; h->k_count[sp.kid]++; }
        h->free_events.push_back(sp.a); h->free_events.push_back(sp.b);
    }
    h->spans.clear();
}

bool supported_n(int n) { return n == 32 || n == 64 || n == 128 || n == 256 || n == 512 || n == 1024 || n == 2048; }

Geom geom_of(const wfm_model* h) {
    Geom g;
    g.N = h->N; g.nz_global = h->nz_global; g.z0 = h->z0; g.nzl = h->nzl; g.nzm = h->nzm; g.dz = h->dz;
    g.psf_norm = 1.0 / ((double)h->N * (double)h->N * (double)h->nz_global);   // WFM:284
    return g;
}

template <typename T> int col_tile(int N) {
    switch (N) {
        case 32: return PipeCfg<T, 32>::C;
        case 64: return PipeCfg<T, 64>::C;
        case 128: return PipeCfg<T, 128>::C;
        case 256: return PipeCfg<T, 256>::C;
        case 512: return PipeCfg<T, 512>::C;
        case 1024: return PipeCfg<T, 1024>::C;
        default: return PipeCfg<T, 2048>::C;
    }
}

int upload_twiddles(wfm_model* h) {
    const int N = h->N;
    {
        std::vector<double2> e(WFM_CIS_ENTRIES);
        for (int k = 0; k < WFM_CIS_ENTRIES; ++k) {
            long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)WFM_CIS_ENTRIES;
            e[k].x = (double)cosl(a); e[k].y = (double)sinl(a);
        }
        WFM_CK(h, h->cis_tab.ensure(sizeof(double2) * WFM_CIS_ENTRIES));
        WFM_CK(h, cudaMemcpy(h->cis_tab.p, e.data(), sizeof(double2) * WFM_CIS_ENTRIES, cudaMemcpyHostToDevice));
    }
    if (h->generic) {
        std::vector<double2> t(N);
        for (int m = 0; m < N; ++m) {
            long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)N;
            t[m].x = (double)cosl(a); t[m].y = (double)(-sinl(a));
        }
        WFM_CK(h, h->tw64.ensure(sizeof(double2) * N));
        WFM_CK(h, cudaMemcpy(h->tw64.p, t.data(), sizeof(double2) * N, cudaMemcpyHostToDevice));
        return WFM_OK;
    }
    if (h->precision == WFM_F64) {
        std::vector<double2> t(N);
        for (int m = 0; m < N; ++m) {
            long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)N;
            t[m].x = (double)cosl(a); t[m].y = (double)(-sinl(a));
        }
        WFM_CK(h, h->tw.ensure(sizeof(double2) * N));
        WFM_CK(h, cudaMemcpy(h->tw.p, t.data(), sizeof(double2) * N, cudaMemcpyHostToDevice));
    } else {
        std::vector<float2> t(N);
        for (int m = 0; m < N; ++m) {
            long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)N;
            t[m].x = (float)cosl(a); t[m].y = (float)(-sinl(a));
        }
        WFM_CK(h, h->tw.ensure(sizeof(float2) * N));
        WFM_CK(h, cudaMemcpy(h->tw.p, t.data(), sizeof(float2) * N, cudaMemcpyHostToDevice));
    }
    return WFM_OK;
}

// Rows / columns of the pupil plane that can hold a non-zero value: union of mapPupil, the
// support of the basis and anything loaded through the escape hatch.
int rebuild_activity(wfm_model* h) {
    if (!h->activity_dirty) return WFM_OK;
    const int N = h->N, npix = h->npix();
    std::vector<uint8_t> sup(npix, 0);
    for (int i = 0; i < npix; ++i)
        sup[i] = (uint8_t)((h->h_map.empty() ? 0 : h->h_map[i]) | (h->h_zsup.empty() ? 0 : h->h_zsup[i]) |
                           (h->h_esc.empty() ? 0 : h->h_esc[i]));
    std::vector<int> ax, ay, ix(N, -1), iy(N, -1);
    std::vector<uint8_t> cx_any(N, 0), cy_any(N, 0);
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x)
            if (sup[x + N * y]) { cx_any[x] = 1; cy_any[y] = 1; }
    for (int x = 0; x < N; ++x) if (cx_any[x]) { ix[x] = (int)ax.size(); ax.push_back(x); }
    for (int y = 0; y < N; ++y) if (cy_any[y]) { iy[y] = (int)ay.size(); ay.push_back(y); }
    if (ax.empty()) { ix[0] = 0; ax.push_back(0); }
    if (ay.empty()) { iy[0] = 0; ay.push_back(0); }
    h->nax = (int)ax.size(); h->nay = (int)ay.size();
    h->narrow = getenv("WFM_NO_NARROW") == nullptr;
    for (int x : ax) if (x >= N / 4 && x < N - N / 4) h->narrow = false;
    for (int y : ay) if (y >= N / 4 && y < N - N / 4) h->narrow = false;
    const int C = h->generic ? 4 : (h->precision == WFM_F64 ? col_tile<double>(N) : col_tile<float>(N));
    h->pitch = (h->nax + C - 1) / C * C;
    h->ctile = C;
    WFM_CK(h, h->act_x.ensure(sizeof(int) * ax.size()));
    WFM_CK(h, h->act_y.ensure(sizeof(int) * ay.size()));
    WFM_CK(h, h->inv_x.ensure(sizeof(int) * N));
    WFM_CK(h, h->inv_y.ensure(sizeof(int) * N));
    WFM_CK(h, h->support.ensure(npix));
    std::vector<int> cells, ins;                  // support cells of the compact strip and their pixel indices
    for (int tile = 0; tile * C < h->nax; ++tile)          // tile-major order == memory order of the strip
        for (int ky = 0; ky < N; ++ky)
            for (int c = 0; c < C && tile * C + c < h->nax; ++c)
                if (sup[ax[tile * C + c] + N * ky]) { cells.push_back((tile * N + ky) * C + c); ins.push_back(ax[tile * C + c] + N * ky); }
    h->ncells = (int)cells.size();
    WFM_CK(h, h->cell_list.ensure(sizeof(int) * (cells.size() + 1)));
    WFM_CK(h, h->in_list.ensure(sizeof(int) * (cells.size() + 1)));
    h->basis_packed = false;
    WFM_CK(h, cudaStreamSynchronize(h->stream));
    if (!cells.empty())
        WFM_CK(h, cudaMemcpy(h->cell_list.p, cells.data(), sizeof(int) * cells.size(), cudaMemcpyHostToDevice));
    if (!ins.empty())
        WFM_CK(h, cudaMemcpy(h->in_list.p, ins.data(), sizeof(int) * ins.size(), cudaMemcpyHostToDevice));
    WFM_CK(h, cudaMemcpy(h->act_x.p, ax.data(), sizeof(int) * ax.size(), cudaMemcpyHostToDevice));
    WFM_CK(h, cudaMemcpy(h->act_y.p, ay.data(), sizeof(int) * ay.size(), cudaMemcpyHostToDevice));
    WFM_CK(h, cudaMemcpy(h->inv_x.p, ix.data(), sizeof(int) * N, cudaMemcpyHostToDevice));
    WFM_CK(h, cudaMemcpy(h->inv_y.p, iy.data(), sizeof(int) * N, cudaMemcpyHostToDevice));
    WFM_CK(h, cudaMemcpy(h->support.p, sup.data(), npix, cudaMemcpyHostToDevice));
    h->activity_dirty = false;
    return WFM_OK;
}

template <class K> int set_smem(wfm_model* h, K kfn, size_t bytes) {
    if (bytes > 48 * 1024) WFM_CK(h, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return WFM_OK;
}

// Pupil arrays -> strip layout (after any setter, before the next pipeline launch).
int pack_strip(wfm_model* h) {
    if (!h->strip_dirty) return WFM_OK;
    const size_t cells = (size_t)h->N * h->pitch, all = cells * h->nbatch;
    WFM_CK(h, h->s_rho.ensure(8 * all)); WFM_CK(h, h->s_phi.ensure(8 * all));
    WFM_CK(h, h->s_psi.ensure(8 * all)); WFM_CK(h, h->s_flags.ensure(all));
    KernelSpan span(h, WFM_K_SETTERS);
    auto kfn = &k_pack_strip;
    WFM_LAUNCH_PDL(kfn, dim3((unsigned)((cells + 255) / 256), (unsigned)h->nbatch), dim3(256), 0, h->stream, (double*)h->s_rho.p,
               (double*)h->s_phi.p, (double*)h->s_psi.p, (uint8_t*)h->s_flags.p, (const double*)h->rho.p,
               (const double*)h->phi.p, (const double*)h->psi.p, (const uint8_t*)h->mask.p,
               (const uint8_t*)h->support.p, (const int*)h->act_x.p, h->N, h->nax, h->pitch, h->ctile);
    WFM_CK_LAUNCH(h, "k_pack_strip");
    h->strip_dirty = false;
    return WFM_OK;
}
Strip strip_of(const wfm_model* h) {
    Strip st;
    st.rho = (const double*)h->s_rho.p; st.phi = (const double*)h->s_phi.p; st.psi = (const double*)h->s_psi.p;
    st.flags = (const uint8_t*)h->s_flags.p;
    return st;
}

// ---- pipeline control ---------------------------------------------------------------------------------
// Ring / lag sizing: LAG ~ 1.5x the planes whose A-items are in flight at once, RING = 2*LAG + 2, with the
// ring kept within ~48 MB so that it stays L2-resident (126 MB L2 shared with the streaming traffic).
struct PipePlan { int ring, lag, nA, nB, grid; };

PipePlan plan_pipeline(const wfm_model* h, int nA, int nB, size_t plane_bytes, int ctas_per_sm, int planes) {
    PipePlan p;
    p.nA = nA; p.nB = nB;
    if (const char* e = getenv("WFM_PIPE_CTAS")) { int v = atoi(e); if (v >= 1 && v < ctas_per_sm) ctas_per_sm = v; }   // experiment knob
    const int nctas = h->num_sms * ctas_per_sm;
    // A CTA holds its current item plus, for the last part of it, the next one: ~1.3*nctas consecutive queue
    // entries are in flight, and B(p) must be queued about twice that far behind A(p) or its CTA stalls on
    // cntA[p].  Measured at 512^2 fp64 (ms/step): lag 8: 1.14, 12: 1.00, 16: 0.949, 20: 0.940, 26: 0.963.
    int lag = (13 * nctas / 5 + (nA + nB) - 1) / (nA + nB);
    if (lag < 2) lag = 2;
    size_t budget = 64u << 20;                 // ring = 2*lag+2 planes, meant to stay L2-resident (126 MB L2)
    if (const char* e = getenv("WFM_PIPE_RING_MB")) { int v = atoi(e); if (v > 0) budget = (size_t)v << 20; }
    if (const char* e = getenv("WFM_PIPE_LAG")) { int v = atoi(e); if (v >= 1) lag = v; }
    while (lag > 2 && (size_t)(2 * lag + 2) * plane_bytes > budget) --lag;
    int ring = 2 * lag + 2;
    if (ring > planes) ring = planes;
    if (lag >= ring) lag = ring > 1 ? ring - 1 : 1;
    p.ring = ring; p.lag = lag;
    const long items = (long)planes * (nA + nB);
    p.grid = (int)(items < nctas ? items : nctas);
    return p;
}

int prepare_ctl(wfm_model* h, const PipePlan& pp, PipeCtl& ctl, int roles) {
    const size_t words = 4 + 2 * (size_t)h->nzl;
    if (h->ctl.bytes < sizeof(unsigned) * words || h->ctl_dirty) {      // the kernels leave the block clean
        WFM_CK(h, h->ctl.ensure(sizeof(unsigned) * words));
        WFM_CK(h, cudaMemsetAsync(h->ctl.p, 0, sizeof(unsigned) * words, h->stream));
        h->ctl_dirty = false;
    }
    unsigned* base = (unsigned*)h->ctl.p;
    ctl.queue = base; ctl.err = base + 1; ctl.done = base + 2; ctl.cntA = base + 4; ctl.cntB = base + 4 + h->nzl;
    ctl.ring = pp.ring; ctl.lag = pp.lag; ctl.nA = pp.nA; ctl.nB = pp.nB; ctl.roles = roles; ctl.nzm = h->nzm;
    ctl.ring_stride = h->N * h->pitch;
    return WFM_OK;
}

int pipe_roles() {
    const char* e = getenv("WFM_PIPE_ROLES");      // debugging / profiling aid: 1 = A only, 2 = B only
    int r = e ? atoi(e) : 3;
    return (r >= 1 && r <= 3) ? r : 3;
}

// ---- computePsf ---------------------------------------------------------------------------------
template <typename T, int N> int launch_psf(wfm_model* h) {
    using Cfg = PipeCfg<T, N>;
    auto kfn = &k_psf_pipeline<T, N, false>;
    if constexpr (Plan<N>::R1 == 8 || Plan<N>::R1 == 16) {
        if (h->narrow) kfn = &k_psf_pipeline<T, N, true>;       // legs 2..5 of the first stages are zero: pruned kernels
    }
    int rc = set_smem(h, kfn, Cfg::SMEM); if (rc) return rc;
    rc = pack_strip(h); if (rc) return rc;
    const int nA = h->pitch / Cfg::C, nB = N / Cfg::ROWS_PER_ITEM;
    const size_t plane_bytes = sizeof(cx<T>) * (size_t)N * h->pitch;
    // plane window of this launch (chunked host paths; single-model handles only)
    const int P0 = h->winN ? h->win0 : 0, PN = h->winN ? h->winN : h->nzl;
    const PipePlan pp = plan_pipeline(h, nA, nB, plane_bytes, Cfg::MINB, PN);
    WFM_CK(h, h->scratch.ensure(plane_bytes * pp.ring));
    PsfArgs<T> a;
    a.g = geom_of(h);
    if (h->winN) { a.g.z0 += P0; a.g.nzl = PN; a.g.nzm = PN; }
    a.st = strip_of(h);
    a.inv_x = (const int*)h->inv_x.p;
    a.nax = h->nax; a.pitch = h->pitch;
    a.tw = (const cx<T>*)h->tw.p;
    a.cis = (const double2*)h->cis_tab.p;
    a.T1 = (cx<T>*)h->scratch.p;
    a.cpx = (cx<T>*)h->cpx.p + (size_t)P0 * N * N; a.psf = (T*)h->psf.p + (size_t)P0 * N * N;
    PipeCtl ctl;
    rc = prepare_ctl(h, pp, ctl, pipe_roles()); if (rc) return rc;
    if (h->copy_pending) WFM_CK(h, cudaStreamWaitEvent(h->stream, h->ev_copied, 0));   // psf is still being read out
    KernelSpan span(h, WFM_K_PSF);
    WFM_LAUNCH_PDL(kfn, dim3(pp.grid), dim3(Cfg::THREADS), Cfg::SMEM, h->stream, a, ctl);
    WFM_CK_LAUNCH(h, "k_psf_pipeline");
    h->pipe_checks++;
    return WFM_OK;
}

// z-sums and basis contractions of the per-plane integrands (shared by the pipelines and the any-N path)
int launch_jac_reduce(wfm_model* h, unsigned kinds, const Geom& g, const double* Gj, const double* Gm, int last_plane_only,
                      double* grad_dev) {
    {
        KernelSpan span(h, WFM_K_JAC_REDUCE);
        if (!h->basis_packed && h->nzern > 0 && h->ncells > 0) {
            WFM_CK(h, h->Zs.ensure(sizeof(double) * (size_t)h->nzern * h->ncells));
            auto kpk = &k_pack_basis;
            WFM_LAUNCH(kpk, dim3((h->ncells + 255) / 256), dim3(256), 0, h->stream, (double*)h->Zs.p,
                       (const double*)h->Z.p, (const int*)h->in_list.p, h->ncells, h->nzern, h->npix());
            WFM_CK_LAUNCH(h, "k_pack_basis");
            h->basis_packed = true;
        }
        ReduceArgs r;
        r.g = g; r.Gj = Gj; r.Gm = Gm; r.pitch = h->pitch;
        r.Zs = (const double*)h->Zs.p; r.psi = (const double*)h->psi.p; r.flags = (const uint8_t*)h->s_flags.p;
        r.cell_list = (const int*)h->cell_list.p; r.in_list = (const int*)h->in_list.p; r.ncells = h->ncells;
        r.nphase = h->nphase; r.nmod = h->nmod; r.phase_off = h->radial ? 1 : 3;
        r.kinds = kinds; r.last_plane_only = last_plane_only;
        r.dxy = h->dxy; r.lambda_ni = h->lambda_ni; r.deltaX = h->deltaX; r.deltaY = h->deltaY;
        r.glen = h->glen();
        const bool batch = h->nbatch > 1;
        r.bpar = batch ? (const double*)h->bpar_dev.p : nullptr;
        const int nblocks = h->ncells > 0 ? (h->ncells + WFM_RED_THREADS - 1) / WFM_RED_THREADS : 1;
        // the longest chunks that still give every SM two CTAs (512^2 x 256: 64-plane chunks, one wave of 360 CTAs;
        // 256^2 x 128: 16-plane chunks, 184 CTAs instead of 46)
        r.batches = WFM_RED_BATCHES;
        while (r.batches > 1 &&
               (long)nblocks * ((h->nzm + WFM_RED_PLANES * r.batches - 1) / (WFM_RED_PLANES * r.batches)) * h->nbatch < 2L * h->num_sms)
            r.batches >>= 1;
        r.cpm = (h->nzm + WFM_RED_PLANES * r.batches - 1) / (WFM_RED_PLANES * r.batches);
        const int nchunks = r.cpm * h->nbatch;
        WFM_CK(h, h->block_part.ensure(sizeof(double) * (size_t)nblocks * nchunks * r.glen));
        r.block_part = (double*)h->block_part.p;
        auto kred = &k_jac_reduce;
        WFM_LAUNCH_PDL(kred, dim3(nblocks, nchunks), dim3(WFM_RED_THREADS), 0, h->stream, r);
        WFM_CK_LAUNCH(h, "k_jac_reduce");
        double nbeta = 0.0;
        if (h->nmod > 0) {
            double s = 0.0;
            for (int k = 0; k < h->nmod; ++k) s += h->beta.v[k] * h->beta.v[k];
            nbeta = 1.0 / std::sqrt(s);                                        // WFM:435
        }
        XchgArgs xc;
        memset(&xc, 0, sizeof(xc));
        if (h->xchg && h->xchg->connected) {                 // sum over the ranks inside k_jac_final (peer memory)
            Exchange* x = h->xchg;
            if (batch) return h->fail(WFM_ERR_UNSUPPORTED, "the peer-memory gradient exchange needs a single-model handle");
            if (r.glen > x->glen_cap) return h->fail(WFM_ERR_STATE, "gradient vector longer than the exchange buffer: reconnect");
            xc.world = x->world; xc.rank = x->rank; xc.epoch = ++x->epoch;
            for (int k = 0; k < x->world; ++k) {
                xc.slots[k] = (double*)x->mapped[k];
                xc.flags[k] = (unsigned*)((char*)x->mapped[k] + x->slot_bytes);
            }
            xc.ticket = (unsigned*)x->local.p; xc.err = (unsigned*)x->local.p + 1;
        }
        auto kfin = &k_jac_final;
        WFM_LAUNCH_PDL(kfin, dim3(r.glen, h->nbatch), dim3(WFM_FINAL_THREADS), 0, h->stream, (const double*)r.block_part,
                   nblocks * r.cpm, r.glen, h->nphase, g.psf_norm, h->beta, nbeta, kinds,
                   batch ? (const double*)h->beta_dev.p : (const double*)nullptr, h->nmod, r.bpar, grad_dev, xc);
        WFM_CK_LAUNCH(h, "k_jac_final");
    }
    return WFM_OK;
}

// ---- apply_J_* ------------------------------------------------------------------------------------
template <typename T, int N> int launch_jac(wfm_model* h, unsigned kinds, const void* q_dev, double* grad_dev) {
    using Cfg = PipeCfg<T, N>;
    auto kfn = &k_jac_pipeline<T, N, false>;
    int ctas_per_sm = Cfg::template minb_jac<false>();
    if constexpr (Plan<N>::RL >= 4) {
        if (h->narrow) {                                        // outputs outside the legs that can hit the support are skipped
            kfn = &k_jac_pipeline<T, N, true>;
            ctas_per_sm = Cfg::template minb_jac<true>();
        }
    }
    int rc = set_smem(h, kfn, Cfg::SMEM_JAC); if (rc) return rc;
    rc = pack_strip(h); if (rc) return rc;
    const int nA = N / Cfg::ROWS_PER_ITEM_JAC, nB = h->pitch / Cfg::C;
    const size_t plane_bytes = sizeof(cx<T>) * (size_t)N * h->pitch;
    // plane window of this launch (chunked host path): the z-sums run once, after the last window (win0 + winN == nzl)
    const int P0 = h->winN ? h->win0 : 0, PN = h->winN ? h->winN : h->nzl;
    const PipePlan pp = plan_pipeline(h, nA, nB, plane_bytes, ctas_per_sm, PN);
    WFM_CK(h, h->scratch.ensure(plane_bytes * pp.ring));
    const size_t img = (size_t)N * h->pitch;
    WFM_CK(h, h->Gj.ensure(sizeof(double) * img * h->nzl));
    if (kinds & WFM_J_MODULUS) WFM_CK(h, h->Gm.ensure(sizeof(double) * img * h->nzl));
    JacArgs<T> a;
    a.g = geom_of(h);
    if (h->winN) { a.g.z0 += P0; a.g.nzl = PN; a.g.nzm = PN; }
    a.cpx = (const cx<T>*)h->cpx.p + (size_t)P0 * N * N; a.q = (const T*)q_dev + (size_t)P0 * N * N;
    a.st = strip_of(h);
    a.inv_x = (const int*)h->inv_x.p; a.nax = h->nax; a.pitch = h->pitch;
    a.tw = (const cx<T>*)h->tw.p;
    a.cis = (const double2*)h->cis_tab.p;
    a.T2 = (cx<T>*)h->scratch.p;
    a.Gj = (double*)h->Gj.p + (size_t)P0 * img;
    a.Gm = (kinds & WFM_J_MODULUS) ? (double*)h->Gm.p + (size_t)P0 * img : nullptr;
    a.last_plane_only = (h->modulus_mode == WFM_MODULUS_REFERENCE_LAST_PLANE) ? 1 : 0;
    {
        PipeCtl ctl;
        rc = prepare_ctl(h, pp, ctl, pipe_roles()); if (rc) return rc;
        KernelSpan span(h, WFM_K_JAC);
        WFM_LAUNCH_PDL(kfn, dim3(pp.grid), dim3(Cfg::THREADS), Cfg::SMEM_JAC, h->stream, a, ctl);
        WFM_CK_LAUNCH(h, "k_jac_pipeline");
        h->pipe_checks++;
    }
    if (h->winN && P0 + PN < h->nzl) return WFM_OK;          // more windows to come
    return launch_jac_reduce(h, kinds, geom_of(h), (const double*)h->Gj.p,
                             (kinds & WFM_J_MODULUS) ? (const double*)h->Gm.p : nullptr, a.last_plane_only, grad_dev);
}

#ifdef WFM_ONLY_512
#define WFM_ONLY_N 512
#endif
#ifdef WFM_ONLY_N   /* quick experiment builds: instantiate one size only */
#define WFM_DISPATCH_N(FN, h, ...)                                                     \
    switch ((h)->N) {                                                                  \
        case WFM_ONLY_N: return FN<T, WFM_ONLY_N>(h, ##__VA_ARGS__);                   \
        default: return (h)->fail(WFM_ERR_UNSUPPORTED, "unsupported N=%d", (h)->N);    \
    }
#else
#define WFM_DISPATCH_N(FN, h, ...)                                                     \
    switch ((h)->N) {                                                                  \
        case 32: return FN<T, 32>(h, ##__VA_ARGS__);                                   \
        case 64: return FN<T, 64>(h, ##__VA_ARGS__);                                   \
        case 128: return FN<T, 128>(h, ##__VA_ARGS__);                                 \
        case 256: return FN<T, 256>(h, ##__VA_ARGS__);                                 \
        case 512: return FN<T, 512>(h, ##__VA_ARGS__);                                 \
        case 1024: return FN<T, 1024>(h, ##__VA_ARGS__);                               \
        case 2048: return FN<T, 2048>(h, ##__VA_ARGS__);                               \
        default: return (h)->fail(WFM_ERR_UNSUPPORTED, "unsupported N=%d", (h)->N);    \
    }
#endif

// ---- any-N path (wfm_generic.cuh): the same pruned row-column transform as plain DFT sums ------------------
struct GenPlan { int chunk; size_t s_elems, t_elems; };
GenPlan plan_generic(const wfm_model* h) {
    GenPlan p;
    p.s_elems = (size_t)h->nay * h->nax;                 // active rows x columns
    p.t_elems = (size_t)h->N * h->nax;                   // all rows x active columns
    const size_t per_plane = 16 * (p.s_elems + p.t_elems);
    size_t c = (256u << 20) / (per_plane ? per_plane : 1);
    if (c < 1) c = 1;
    if (c > (size_t)h->nzl) c = (size_t)h->nzl;
    if (c > 4096) c = 4096;
    p.chunk = (int)c;
    return p;
}
dim3 gen_grid(int nx, int ny, int nz) { return dim3((nx + WFM_GEN_BX - 1) / WFM_GEN_BX, (ny + WFM_GEN_BY - 1) / WFM_GEN_BY, nz); }

template <typename T> int launch_psf_generic(wfm_model* h) {
    const GenPlan gp = plan_generic(h);
    WFM_CK(h, h->scratch.ensure(16 * (gp.s_elems + gp.t_elems) * gp.chunk));
    double2* S = (double2*)h->scratch.p;
    double2* Tm = S + gp.s_elems * gp.chunk;
    const Geom g = geom_of(h);
    const dim3 blk(WFM_GEN_BX, WFM_GEN_BY);
    if (h->copy_pending) WFM_CK(h, cudaStreamWaitEvent(h->stream, h->ev_copied, 0));
    KernelSpan span(h, WFM_K_PSF);
    for (int p0 = 0; p0 < h->nzl; p0 += gp.chunk) {
        const int nb = std::min(gp.chunk, h->nzl - p0);
        auto k1 = &k_gen_synth;
        WFM_LAUNCH(k1, gen_grid(h->nax, h->nay, nb), blk, 0, h->stream, S, (const double*)h->rho.p, (const double*)h->phi.p,
                   (const double*)h->psi.p, (const int*)h->act_x.p, (const int*)h->act_y.p, h->nax, h->nay, g, p0,
                   h->precision == WFM_F32 ? 1 : 0);
        auto k2 = &k_gen_dft_slow;
        WFM_LAUNCH(k2, gen_grid(h->nax, h->N, nb), blk, 0, h->stream, Tm, (const double2*)S, (const double2*)h->tw64.p, h->N,
                   h->nay, (const int*)h->act_y.p, h->N, (const int*)nullptr, h->nax);
        auto k3 = &k_gen_psf_rows<T>;
        WFM_LAUNCH(k3, gen_grid(h->N, h->N, nb), blk, 0, h->stream, (cx<T>*)h->cpx.p, (T*)h->psf.p, (const double2*)Tm,
                   (const double2*)h->tw64.p, (const int*)h->act_x.p, h->nax, g, p0);
    }
    WFM_CK_LAUNCH(h, "any-N PSF kernels");
    return WFM_OK;
}

template <typename T> int launch_jac_generic(wfm_model* h, unsigned kinds, const void* q_dev, double* grad_dev) {
    int rc = pack_strip(h); if (rc) return rc;
    const GenPlan gp = plan_generic(h);
    WFM_CK(h, h->scratch.ensure(16 * (gp.s_elems + gp.t_elems) * gp.chunk));
    double2* B = (double2*)h->scratch.p;
    double2* U = B + gp.s_elems * gp.chunk;
    const size_t img = (size_t)h->N * h->pitch;
    WFM_CK(h, h->Gj.ensure(sizeof(double) * img * h->nzl));
    if (kinds & WFM_J_MODULUS) WFM_CK(h, h->Gm.ensure(sizeof(double) * img * h->nzl));
    const Geom g = geom_of(h);
    const int last_plane_only = (h->modulus_mode == WFM_MODULUS_REFERENCE_LAST_PLANE) ? 1 : 0;
    double* Gm = (kinds & WFM_J_MODULUS) ? (double*)h->Gm.p : nullptr;
    const dim3 blk(WFM_GEN_BX, WFM_GEN_BY);
    {
        KernelSpan span(h, WFM_K_JAC);
        for (int p0 = 0; p0 < h->nzl; p0 += gp.chunk) {
            const int nb = std::min(gp.chunk, h->nzl - p0);
            auto k1 = &k_gen_jac_rows<T>;
            WFM_LAUNCH(k1, gen_grid(h->nax, h->N, nb), blk, 0, h->stream, U, (const cx<T>*)h->cpx.p, (const T*)q_dev,
                       (const double2*)h->tw64.p, (const int*)h->act_x.p, h->nax, g, p0);
            auto k2 = &k_gen_dft_slow;
            WFM_LAUNCH(k2, gen_grid(h->nax, h->nay, nb), blk, 0, h->stream, B, (const double2*)U, (const double2*)h->tw64.p,
                       h->N, h->N, (const int*)nullptr, h->nay, (const int*)h->act_y.p, h->nax);
            auto k3 = &k_gen_jac_trig;
            WFM_LAUNCH(k3, gen_grid(h->nax, h->nay, nb), blk, 0, h->stream, (double*)h->Gj.p, Gm, (const double2*)B, strip_of(h),
                       (const int*)h->act_y.p, h->nax, h->nay, h->pitch, h->ctile, g, p0, last_plane_only,
                       h->precision == WFM_F32 ? 1 : 0);
        }
        WFM_CK_LAUNCH(h, "any-N Jacobian kernels");
    }
    return launch_jac_reduce(h, kinds, g, (const double*)h->Gj.p, Gm, last_plane_only, grad_dev);
}

template <typename T> int dispatch_psf(wfm_model* h) {
    if (h->generic) return launch_psf_generic<T>(h);
    WFM_DISPATCH_N(launch_psf, h)
}
template <typename T> int dispatch_jac(wfm_model* h, unsigned kinds, const void* q, double* g) {
    if (h->generic) return launch_jac_generic<T>(h, kinds, q, g);
    WFM_DISPATCH_N(launch_jac, h, kinds, q, g)
}

// Planes per chunk of the host-buffer paths (wfm_get_psf_async, wfm_apply_j_* with a host q); 0 = one piece.
// A chunk is a plane window of the pipelines: chunk c's kernel runs while chunk c+1 (q, host -> device) or chunk c-1
// (psf, device -> host) is on the PCIe link, so that only one chunk's kernel time is left outside the copies.
int host_chunk_planes(const wfm_model* h) {
    if (h->generic || h->nbatch > 1 || h->multi()) return 0;
    int nch = 4;                                  // (2 .. 8 chunks measure alike, 11.1-11.8 ms per 2 x 537 MB step; 1: 12.4, 16: 13.1)
    if (const char* e = getenv("WFM_HOST_CHUNKS")) nch = atoi(e);
    if (nch <= 1) return 0;
    const size_t plane = (size_t)h->npix() * h->esz();
    int cp = (h->nzl + nch - 1) / nch;
    size_t min_bytes = (size_t)32 << 20;                                         // at least 32 MB per copy
    if (const char* e = getenv("WFM_HOST_CHUNK_MIN_BYTES")) min_bytes = (size_t)atoll(e);   // (tests: chunk small stacks too)
    const int min_planes = (int)((min_bytes + plane - 1) / plane) > 0 ? (int)((min_bytes + plane - 1) / plane) : 1;
    if (cp < min_planes) cp = min_planes;
    if (cp * 2 > h->nzl) return 0;
    return cp;
}

int ensure_chunk_events(wfm_model* h, std::vector<cudaEvent_t>& ev, int n) {
    while ((int)ev.size() < n) {
        cudaEvent_t e = nullptr;
        WFM_CK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ev.push_back(e);
    }
    return WFM_OK;
}

// computePsf().  cp > 0: in plane windows of cp planes; after_window(p0, np) is called right after window [p0, p0+np)
// has been queued on the handle's stream (the asynchronous read-back queues that window's device -> host copy there).
template <class F> int compute_psf_windows(wfm_model* h, int cp, F after_window) {
    if (!h->have_rho) return h->fail(WFM_ERR_STATE, "pupil modulus not set: call wfm_set_modulus or wfm_set_pupil_arrays first");
    WFM_ENTER(h);
    int rc = rebuild_activity(h); if (rc) return rc;
    const size_t vox = (size_t)h->npix() * h->nzl;
    WFM_CK(h, h->cpx.ensure(vox * 2 * h->esz()));
    WFM_CK(h, h->psf.ensure(vox * h->esz()));
    if (cp <= 0) cp = h->nzl;
    for (int p0 = 0; p0 < h->nzl; p0 += cp) {
        const int np = std::min(cp, h->nzl - p0);
        if (np < h->nzl) { h->win0 = p0; h->winN = np; }
        rc = (h->precision == WFM_F64) ? dispatch_psf<double>(h) : dispatch_psf<float>(h);
        h->win0 = 0; h->winN = 0;
        if (rc) return rc;
        rc = after_window(p0, np); if (rc) return rc;
    }
    h->pstate = 1;                                                             // WFM:395
    return WFM_OK;
}

int compute_psf_impl(wfm_model* h) {
    if (h->pstate > 0) return WFM_OK;                                          // WFM:207
    return compute_psf_windows(h, 0, [](int, int) { return WFM_OK; });
}

// a setter changed a pupil array: the PSF is stale (freeMem(), WFM:1970-1974) and so is the packed strip
int invalidate(wfm_model* h) { h->pstate = 0; h->strip_dirty = true; return WFM_OK; }
// freeMem() alone: the PSF is stale, the pupil arrays (and their strip copy) are what they were
int invalidate_psf(wfm_model* h) { h->pstate = 0; return WFM_OK; }

// After a synchronisation point: did a pipeline dependency wait time out?  (It cannot by
// construction; the flag turns a would-be hang into an error code.)
int check_pipeline(wfm_model* h) {
    if (!h->pipe_checks || !h->ctl.p) return WFM_OK;
    h->pipe_checks = 0;
    unsigned flag = 0;
    WFM_CK(h, cudaMemcpy(&flag, (unsigned*)h->ctl.p + 1, sizeof(unsigned), cudaMemcpyDeviceToHost));
    if (flag) { h->ctl_dirty = true; return h->fail(WFM_ERR_INTERNAL, "pipeline dependency wait timed out"); }
    return WFM_OK;
}                 // WFM:1970-1974

int elementwise_grid(int n) { return (n + 255) / 256; }

// A new basis invalidates coefficient vectors that no longer fit it: they are dropped (nPhase / nModulus = 0, phi
// cleared) exactly as setNPhase / setNModulus rebuild their spaces with the basis (WFM:1899-1914, 1939-1961).
int revalidate_coefs(wfm_model* h) {
    const int off = h->radial ? 1 : 3;
    if (h->nphase > 0 && h->nphase + off > h->nzern) {
        h->nphase = 0; h->alpha_b.clear();
        WFM_CK(h, cudaMemsetAsync(h->phi.p, 0, 8 * (size_t)h->npix() * h->nbatch, h->stream));
    }
    if (h->nmod > h->nzern) { h->nmod = 0; h->beta_b.clear(); h->have_rho = false; }
    return WFM_OK;
}

int upload_table(wfm_model* h, DevBuf& buf, const std::vector<double>& v) {
    WFM_CK(h, buf.ensure(8 * v.size()));
    WFM_CK(h, cudaMemcpyAsync(buf.p, v.data(), 8 * v.size(), cudaMemcpyHostToDevice, h->stream));
    return WFM_OK;
}
int upload_bpar(wfm_model* h) { return upload_table(h, h->bpar_dev, h->bpar_h); }

// setPhase / setModulus / setDefocus of every model of a batch handle; tab = [nbatch][n] (stride 0: one row, broadcast)
int batch_set_phase(wfm_model* h, const double* tab, int n, int stride) {
    if (n < 0 || n > WFM_MAX_COEF || (n > 0 && !tab)) return h->fail(WFM_ERR_INVALID_ARG, "bad phase coefficient vector");
    const int off = h->radial ? 1 : 3;
    if (n > 0 && (h->nzern <= 0)) return h->fail(WFM_ERR_STATE, "Zernike basis not set");
    if (n > 0 && n + off > h->nzern)
        return h->fail(WFM_ERR_INVALID_ARG, "phase parameter does not belong to the right space  ");   // WFM:1629
    WFM_ENTER(h);
    h->alpha_b.assign((size_t)h->nbatch * (n > 0 ? n : 1), 0.0);
    for (int b = 0; b < h->nbatch; ++b)
        for (int k = 0; k < n; ++k) h->alpha_b[(size_t)b * n + k] = tab[(size_t)b * stride + k];
    for (int k = 0; k < n; ++k) h->alpha.v[k] = h->alpha_b[k];
    h->nphase = n;
    int rc = upload_table(h, h->alpha_dev, h->alpha_b); if (rc) return rc;
    KernelSpan span(h, WFM_K_SETTERS);
    auto kfn = &k_set_phase_b;
    WFM_LAUNCH(kfn, dim3(elementwise_grid(h->npix()), h->nbatch), dim3(256), 0, h->stream, (double*)h->phi.p,
               (const double*)h->Z.p, (const uint8_t*)h->mask.p, (const double*)h->alpha_dev.p, n, off, h->npix());
    WFM_CK_LAUNCH(h, "k_set_phase_b");
    return invalidate(h);
}

int batch_set_modulus(wfm_model* h, const double* tab, int n, int stride) {
    if (n <= 0 || n > WFM_MAX_COEF || !tab) return h->fail(WFM_ERR_INVALID_ARG, "bad modulus coefficient vector");
    if (h->nzern <= 0) return h->fail(WFM_ERR_STATE, "Zernike basis not set");
    if (n > h->nzern)
        return h->fail(WFM_ERR_INVALID_ARG, "DoubleShapedVector beta does not belong to the modulus space");  // WFM:1592
    WFM_ENTER(h);
    h->beta_b.assign((size_t)h->nbatch * n, 0.0);
    for (int b = 0; b < h->nbatch; ++b) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) { const double v = tab[(size_t)b * stride + k]; h->beta_b[(size_t)b * n + k] = v; s += v * v; }
        h->bpar_h[4 * b + 3] = 1.0 / std::sqrt(s);                             // WFM:1597 (and WFM:435 for the Jacobian)
    }
    for (int k = 0; k < n; ++k) h->beta.v[k] = h->beta_b[k];
    h->nmod = n;
    int rc = upload_table(h, h->beta_dev, h->beta_b); if (rc) return rc;
    rc = upload_bpar(h); if (rc) return rc;
    auto kfn = &k_set_modulus_b;
    WFM_LAUNCH(kfn, dim3(elementwise_grid(h->npix()), h->nbatch), dim3(256), 0, h->stream, (double*)h->rho.p,
               (const double*)h->Z.p, (const uint8_t*)h->mask.p, (const double*)h->beta_dev.p,
               (const double*)h->bpar_dev.p, n, h->npix());
    WFM_CK_LAUNCH(h, "k_set_modulus_b");
    h->have_rho = true;
    return invalidate(h);
}

int batch_set_defocus(wfm_model* h, const double* tab, int n, int stride) {
    if (!tab || (n != 1 && n != 3)) return h->fail(WFM_ERR_INVALID_ARG, "bad defocus  parameters");   // WFM:1530, Q4
    if (!h->have_optics) return h->fail(WFM_ERR_STATE, "optics not set: call wfm_set_optics first");
    WFM_ENTER(h);
    for (int b = 0; b < h->nbatch; ++b) {
        const double* d = tab + (size_t)b * stride;
        if (n == 3) { h->bpar_h[4 * b + 1] = d[1]; h->bpar_h[4 * b + 2] = d[2]; }   // WFM:1518-1520
        h->bpar_h[4 * b] = d[0];                                                     // WFM:1522
    }
    h->lambda_ni = h->bpar_h[0]; h->deltaX = h->bpar_h[1]; h->deltaY = h->bpar_h[2];
    h->ni = h->lambda_ni * h->lambda;
    h->ndefocus = n;
    int rc = upload_bpar(h); if (rc) return rc;
    auto kfn = &k_compute_defocus_b;
    WFM_LAUNCH(kfn, dim3(elementwise_grid(h->npix()), h->nbatch), dim3(256), 0, h->stream, (double*)h->psi.p,
               (uint8_t*)h->mask.p, (const uint8_t*)h->map.p, h->N, h->dxy, (const double*)h->bpar_dev.p);
    WFM_CK_LAUNCH(h, "k_compute_defocus_b");
    return invalidate(h);
}

}  // namespace

// ---- multi-device handles (wfm_create_multi; bodies in wfm_multi.inl) ------------------------------------
namespace wfm_multi {
int destroy(wfm_model* h);
int broadcast(wfm_model* h, const std::function<int(wfm_model*)>& fn);
int sync_meta(wfm_model* h);
int compute_psf(wfm_model* h);
int get_stack(wfm_model* h, void* out, bool cpx, bool async);
int wait_transfers(wfm_model* h);
int apply_host(wfm_model* h, unsigned kinds, const void* q_host, std::vector<double>& g);
int synchronize(wfm_model* h);
int kernel_times(wfm_model* h, double* ms, uint64_t* counts);
int unsupported(wfm_model* h, const char* what);
}  // namespace wfm_multi
#define WFM_MULTI_BCAST(h, call)                                                        \
    do { if ((h)->multi()) {                                                            \
        int rc__ = wfm_multi::broadcast((h), [&](wfm_model* c) { return (call); });     \
        return rc__ ? rc__ : wfm_multi::sync_meta(h); } } while (0)
#define WFM_MULTI_FIRST(h, call)                                                        \
    do { if ((h)->multi()) { wfm_model* c = (h)->parts[0]; int rc__ = (call);           \
        if (rc__) (h)->err = c->err; return rc__; } } while (0)
#define WFM_MULTI_NO(h, what) do { if ((h)->multi()) return wfm_multi::unsupported((h), what); } while (0)

static int jacobian_preconditions(wfm_model* h, unsigned kinds) {
    if (!(kinds & 7u)) return h->fail(WFM_ERR_INVALID_ARG, "no Jacobian selected");
    if ((kinds & WFM_J_PHASE) && h->nphase <= 0) return h->fail(WFM_ERR_STATE, "phase space is empty (nPhase = 0)");
    if ((kinds & WFM_J_MODULUS) && h->nmod <= 0) return h->fail(WFM_ERR_STATE, "modulus coefficients not set");
    if ((kinds & (WFM_J_PHASE | WFM_J_MODULUS)) && h->nzern <= 0) return h->fail(WFM_ERR_STATE, "Zernike basis not set");
    return WFM_OK;
}

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

static int create_impl(wfm_model** out, int nx, int ny, int nz_global, int z0, int nz_local, int nbatch, double dxy,
                       double dz, int precision, int device) {
    if (!out) { g_create_error = "out is NULL"; return WFM_ERR_INVALID_ARG; }
    *out = nullptr;
    if (nx != ny) { g_create_error = "Nx should equal Ny"; return WFM_ERR_INVALID_ARG; }      // WFM:158-160
    if (nx <= 0 || nz_global <= 0 || nz_local <= 0 || z0 < 0 || z0 + nz_local > nz_global) {
        g_create_error = "bad shape / slab"; return WFM_ERR_INVALID_ARG;
    }
    // (the model index is a grid.y dimension of the setters and, times the plane chunks per model, of the reduction)
    if (nbatch < 1 || (long long)nbatch * nz_local > (1ll << 24) ||
        (long long)nbatch * ((nz_local + WFM_RED_PLANES - 1) / WFM_RED_PLANES) > 65535) {
        g_create_error = "bad batch size"; return WFM_ERR_INVALID_ARG;
    }
    if (precision != WFM_F64 && precision != WFM_F32) { g_create_error = "bad precision"; return WFM_ERR_INVALID_ARG; }
    const bool generic = !supported_n(nx);
    if (generic && (nx < 2 || nx > 4096 || nbatch > 1)) {
        g_create_error = nbatch > 1 ? "a batch handle needs Nx to be a power of two in [32, 2048]"
                                    : "Nx must lie in [2, 4096] (powers of two in [32, 2048] take the fast pipelines)";
        return WFM_ERR_UNSUPPORTED;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        g_create_error = "no CUDA device available (this library has no CPU fallback)"; return WFM_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { g_create_error = "bad device index"; return WFM_ERR_INVALID_ARG; }
    wfm_model* h = new (std::nothrow) wfm_model();
    if (!h) { g_create_error = "out of host memory"; return WFM_ERR_NOMEM; }
    h->N = nx; h->nz_global = nz_global; h->z0 = z0; h->nzm = nz_local; h->nbatch = nbatch; h->nzl = nbatch * nz_local;
    h->dxy = dxy; h->dz = dz;
    h->precision = precision; h->device = device; h->generic = generic;
    if (nbatch > 1) h->bpar_h.assign(4 * (size_t)nbatch, 0.0);
    auto bail = [&](int code, const char* what) { g_create_error = what; wfm_destroy(h); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(WFM_ERR_CUDA, "cudaSetDevice failed");
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(WFM_ERR_CUDA, "cudaStreamCreate failed");
    h->stream = h->own_stream;
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) == cudaSuccess && prop.multiProcessorCount > 0)
            h->num_sms = prop.multiProcessorCount;
    }
    const size_t npix = (size_t)nx * nx, all = npix * nbatch;
    // this.phi = new double[Ny*Nx]; this.psi = new double[Ny*Nx]  (WFM:167-168); rho starts empty
    if (h->rho.ensure(8 * all) || h->phi.ensure(8 * all) || h->psi.ensure(8 * all) || h->mask.ensure(all) ||
        h->map.ensure(npix) || h->grad.ensure(8 * (3 + 2 * WFM_MAX_COEF) * (size_t)nbatch))
        return bail(WFM_ERR_NOMEM, "device allocation failed");
    if (cudaMemset(h->rho.p, 0, 8 * all) != cudaSuccess || cudaMemset(h->phi.p, 0, 8 * all) != cudaSuccess ||
        cudaMemset(h->psi.p, 0, 8 * all) != cudaSuccess || cudaMemset(h->mask.p, 0, all) != cudaSuccess ||
        cudaMemset(h->map.p, 0, npix) != cudaSuccess)
        return bail(WFM_ERR_CUDA, "clearing the pupil arrays failed");
    if (upload_twiddles(h) != WFM_OK) return bail(WFM_ERR_CUDA, "twiddle upload failed");
    *out = h;
    return WFM_OK;
}

int wfm_create_slab(wfm_model** out, int nx, int ny, int nz_global, int z0, int nz_local, double dxy,
                    double dz, int precision, int device) {
    return create_impl(out, nx, ny, nz_global, z0, nz_local, 1, dxy, dz, precision, device);
}

int wfm_create_batch(wfm_model** out, int nx, int ny, int nz, int nbatch, double dxy, double dz, int precision,
                     int device) {
    return create_impl(out, nx, ny, nz, 0, nz, nbatch, dxy, dz, precision, device);
}

int wfm_batch_size(const wfm_model* h) { return h ? h->nbatch : 0; }

int wfm_create(wfm_model** out, int nx, int ny, int nz, double dxy, double dz, int precision, int device) {
    return wfm_create_slab(out, nx, ny, nz, 0, nz, dxy, dz, precision, device);
}

int wfm_destroy(wfm_model* h) {
    if (!h) return WFM_OK;
    if (h->multi()) return wfm_multi::destroy(h);
    wfm_exchange_close(h);
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (DevBuf* b : {&h->Z, &h->rho, &h->phi, &h->psi, &h->mask, &h->map, &h->support, &h->act_x, &h->inv_x,
                      &h->act_y, &h->inv_y, &h->cell_list, &h->in_list, &h->Zs, &h->s_rho, &h->s_phi, &h->s_psi, &h->s_flags, &h->tw, &h->cpx, &h->psf, &h->scratch, &h->Gj, &h->Gm, &h->ctl, &h->block_part,
                      &h->grad, &h->qdev, &h->alpha_dev, &h->beta_dev, &h->bpar_dev, &h->tw64, &h->cis_tab})
        b->release();
    drain_spans(h);
    for (cudaEvent_t e : h->free_events) cudaEventDestroy(e);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    if (h->in_stream) { cudaStreamSynchronize(h->in_stream); cudaStreamDestroy(h->in_stream); }
    for (cudaEvent_t e : h->ev_chunk) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_in) cudaEventDestroy(e);
    if (h->stage) { h->stage->release(); delete h->stage; h->stage = nullptr; }
    if (h->ev_ready) cudaEventDestroy(h->ev_ready);
    if (h->ev_copied) cudaEventDestroy(h->ev_copied);
    if (h->ev_order) cudaEventDestroy(h->ev_order);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return WFM_OK;
}

const char* wfm_last_error(const wfm_model* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int wfm_set_stream(wfm_model* h, void* s) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_NO(h, "wfm_set_stream (every device of a multi handle runs its own stream)");
    WFM_ENTER(h);
    cudaStreamSynchronize(h->stream);
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    return WFM_OK;
}

// Ordering against a caller-owned stream without stalling the host (one cached event, re-recorded per call).
static int order_streams(wfm_model* h, cudaStream_t first, cudaStream_t then) {
    if (first == then) return WFM_OK;
    if (!h->ev_order) WFM_CK(h, cudaEventCreateWithFlags(&h->ev_order, cudaEventDisableTiming));
    WFM_CK(h, cudaEventRecord(h->ev_order, first));
    WFM_CK(h, cudaStreamWaitEvent(then, h->ev_order, 0));
    return WFM_OK;
}
int wfm_wait_stream(wfm_model* h, void* s) {          // the handle's stream waits for what `s` holds so far
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_NO(h, "wfm_wait_stream");
    WFM_ENTER(h);
    return order_streams(h, (cudaStream_t)s, h->stream);
}
int wfm_fence_stream(wfm_model* h, void* s) {         // `s` waits for what the handle's stream holds so far
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_NO(h, "wfm_fence_stream");
    WFM_ENTER(h);
    return order_streams(h, h->stream, (cudaStream_t)s);
}

int wfm_synchronize(wfm_model* h) {
    if (!h) return WFM_ERR_INVALID_ARG;
    if (h->multi()) return wfm_multi::synchronize(h);
    WFM_ENTER(h);
    WFM_CK(h, cudaStreamSynchronize(h->stream));
    return check_pipeline(h);
}

int wfm_set_optics(wfm_model* h, double NA, double lambda, double ni) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_BCAST(h, wfm_set_optics(c, NA, lambda, ni));
    if (!(NA > 0) || !(lambda > 0) || !(ni > 0)) return h->fail(WFM_ERR_INVALID_ARG, "NA, lambda and ni must be positive");
    WFM_ENTER(h);
    h->NA = NA; h->lambda = lambda; h->ni = ni;
    h->radius = NA / lambda;                                                   // WFM:165
    h->lambda_ni = ni / lambda;                                                // WFM:166
    h->have_optics = true;
    auto kfn = &k_mask_pupil;
    WFM_LAUNCH(kfn, dim3(elementwise_grid(h->npix())), dim3(256), 0, h->stream, (uint8_t*)h->map.p,
               (uint8_t*)h->mask.p, h->N, h->dxy, h->radius);
    WFM_CK_LAUNCH(h, "k_mask_pupil");
    for (int b = 1; b < h->nbatch; ++b)           // every model of a batch starts from maskPupil = mapPupil
        WFM_CK(h, cudaMemcpyAsync((uint8_t*)h->mask.p + (size_t)b * h->npix(), h->map.p, h->npix(), cudaMemcpyDeviceToDevice, h->stream));
    for (int b = 0; b < h->nbatch && h->nbatch > 1; ++b) h->bpar_h[4 * b] = h->lambda_ni;
    if (h->nbatch > 1) { int rcb = upload_bpar(h); if (rcb) return rcb; }
    h->h_map.resize(h->npix());
    WFM_CK(h, cudaMemcpyAsync(h->h_map.data(), h->map.p, h->npix(), cudaMemcpyDeviceToHost, h->stream));
    WFM_CK(h, cudaStreamSynchronize(h->stream));
    h->activity_dirty = true;
    return invalidate(h);                                                      // WFM:1405
}

int wfm_set_basis(wfm_model* h, const double* Z, int nzern, int radial) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_BCAST(h, wfm_set_basis(c, Z, nzern, radial));
    if (!Z || nzern <= 0) return h->fail(WFM_ERR_INVALID_ARG, "Z is NULL or nzern <= 0");
    WFM_ENTER(h);
    const size_t npix = h->npix();
    WFM_CK(h, cudaStreamSynchronize(h->stream));
    WFM_CK(h, h->Z.ensure(8 * npix * nzern));
    WFM_CK(h, cudaMemcpy(h->Z.p, Z, 8 * npix * nzern, cudaMemcpyHostToDevice));
    h->nzern = nzern; h->radial = radial ? 1 : 0; h->basis_packed = false;
    h->h_zsup.assign(npix, 0);
    for (int k = 0; k < nzern; ++k)
        for (size_t i = 0; i < npix; ++i)
            if (Z[(size_t)k * npix + i] != 0.0) h->h_zsup[i] = 1;
    h->activity_dirty = true;
    { int rcv = revalidate_coefs(h); if (rcv) return rcv; }
    return invalidate(h);
}

// computeZernike() WFM:194-197 on the device.
int wfm_build_basis(wfm_model* h, int nzern, int radial) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_BCAST(h, wfm_build_basis(c, nzern, radial));
    if (nzern <= 0) return h->fail(WFM_ERR_INVALID_ARG, "nzern <= 0");
    if (!h->have_optics) return h->fail(WFM_ERR_STATE, "optics not set: call wfm_set_optics first");
    WFM_ENTER(h);
    const int N = h->N, npix = h->npix();
    // Noll index -> (n, m)  Zernike.java:37-52
    auto noll = [](int J, int& n, int& m) {
        const double n1 = (std::sqrt(1.0 + 8.0 * J) - 1.0) / 2.0;
        n = (int)std::floor(n1);
        if (n1 == (double)n) n -= 1;
        const int k = (n + 1) * (n + 2) / 2;
        m = n - 2 * (int)std::floor((k - J) / 2.0);
    };
    std::vector<ZernMode> modes(nzern);
    memset(modes.data(), 0, sizeof(ZernMode) * nzern);
    int nmax = 0;
    for (int nz = 1; nz < nzern; ++nz) {
        int n, m;
        if (radial) { n = nz; m = 0; } else noll(nz + 1, n, m);
        const int p = (n - m) / 2, q = (n + m) / 2;
        if (p + 1 > WFM_ZERN_MAXS || n > 2 * WFM_ZERN_MAXS - 2)
            return h->fail(WFM_ERR_UNSUPPORTED, "Zernike radial degree %d too high", n);
        ZernMode& md = modes[nz];
        md.n = n; md.m = m;
        md.kind = (m == 0) ? 0 : (((nz + 1) % 2 == 0) ? 1 : 2);                 // Zernike.java:217,241
        md.norm = (m == 0) ? std::sqrt((double)(n + 1)) : std::sqrt((double)(2 * (n + 1)));
        std::vector<double> lfact(n + 1, 0.0);                                 // Zernike.java:75-80
        for (int i = 1; i <= n; ++i) lfact[i] = lfact[i - 1] + std::log((double)i);
        for (int sI = 0; sI <= p; ++sI) {
            double r = std::exp(lfact[n - sI] - lfact[sI] - lfact[p - sI] - lfact[q - sI]);
            md.R[sI] = (sI % 2) ? -r : r;
        }
        if (n > nmax) nmax = n;
    }
    if (nmax < 1) nmax = 1;
    const double radius_px = h->radius * h->dxy * (double)N;                   // WFM:195
    DevBuf dmodes, partial;
    WFM_CK(h, dmodes.ensure(sizeof(ZernMode) * nzern));
    const int nparts = 128;
    WFM_CK(h, partial.ensure(8 * nparts));
    WFM_CK(h, cudaStreamSynchronize(h->stream));
    WFM_CK(h, cudaMemcpy(dmodes.p, modes.data(), sizeof(ZernMode) * nzern, cudaMemcpyHostToDevice));
    WFM_CK(h, h->Z.ensure(8 * (size_t)npix * nzern));
    double* Z = (double*)h->Z.p;
    {
        auto kfn = &k_zernike_modes;
        WFM_LAUNCH(kfn, dim3(elementwise_grid(npix)), dim3(256), 0, h->stream, Z, (const ZernMode*)dmodes.p,
                   nzern, nmax, N, radius_px);
        WFM_CK_LAUNCH(h, "k_zernike_modes");
    }
    auto kdot = &k_dot_partial;
    auto kupd = &k_gs_update;
    auto dot = [&](const double* a, const double* b) {
        WFM_LAUNCH(kdot, dim3(nparts), dim3(WFM_DOT_THREADS), 0, h->stream, a, b, npix, (double*)partial.p);
    };
    auto update = [&](double* zk, const double* zj, int mode) {
        WFM_LAUNCH(kupd, dim3(elementwise_grid(npix)), dim3(256), 0, h->stream, zk, zj,
                   (const double*)partial.p, nparts, npix, mode);
    };
    for (int k = 0; k < nzern; ++k) {              // per-mode L2 normalisation, Zernike.java:156,192,231,255,277
        double* zk = Z + (size_t)k * npix;
        dot(zk, zk); update(zk, zk, 1);
    }
    for (int k = 0; k < nzern; ++k) {              // in-order modified Gram-Schmidt (ASSUMED, WFM:196)
        double* zk = Z + (size_t)k * npix;
        for (int j = 0; j < k; ++j) { dot(Z + (size_t)j * npix, zk); update(zk, Z + (size_t)j * npix, 0); }
        dot(zk, zk); update(zk, zk, 1);
    }
    WFM_CK_LAUNCH(h, "zernike basis kernels");
    WFM_CK(h, cudaStreamSynchronize(h->stream));
    dmodes.release(); partial.release();
    h->nzern = nzern; h->radial = radial ? 1 : 0; h->basis_packed = false;
    h->h_zsup.assign(npix, 0);
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) {
            const double kx = (double)((x > N / 2) ? x - N : x), ky = (double)((y > N / 2) ? y - N : y);
            if (std::sqrt(kx * kx + ky * ky) < radius_px) h->h_zsup[x + N * y] = 1;
        }
    h->activity_dirty = true;
    { int rcv = revalidate_coefs(h); if (rcv) return rcv; }
    return invalidate(h);
}

int wfm_get_basis(wfm_model* h, double* out, int nzern) {
    if (!h || !out) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_FIRST(h, wfm_get_basis(c, out, nzern));
    WFM_ENTER(h);
    if (nzern <= 0 || nzern > h->nzern) return h->fail(WFM_ERR_INVALID_ARG, "nzern out of range");
    WFM_CK(h, cudaStreamSynchronize(h->stream));
    WFM_CK(h, cudaMemcpy(out, h->Z.p, 8 * (size_t)h->npix() * nzern, cudaMemcpyDeviceToHost));
    return WFM_OK;
}

int wfm_set_phase(wfm_model* h, const double* alpha, int n) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_BCAST(h, wfm_set_phase(c, alpha, n));
    if (h->nbatch > 1) return batch_set_phase(h, alpha, n, 0);     // batch handle: the same vector for every model
    if (n < 0 || n > WFM_MAX_COEF || (n > 0 && !alpha)) return h->fail(WFM_ERR_INVALID_ARG, "bad phase coefficient vector");
    const int off = h->radial ? 1 : 3;
    if (n > 0 && (h->nzern <= 0)) return h->fail(WFM_ERR_STATE, "Zernike basis not set");
    if (n > 0 && n + off > h->nzern)
        return h->fail(WFM_ERR_INVALID_ARG, "phase parameter does not belong to the right space  ");   // WFM:1629
    WFM_ENTER(h);
    for (int k = 0; k < n; ++k) h->alpha.v[k] = alpha[k];
    h->nphase = n;
    KernelSpan span(h, WFM_K_SETTERS);
    // the strip is packed and phi is all that changes: support cells only, phi and its strip copy in one pass
    const bool fused = !h->activity_dirty && !h->strip_dirty && h->s_phi.p && h->ncells > 0 && h->phi_clean_off_support &&
                       getenv("WFM_NO_FUSED_SETPHASE") == nullptr;
    if (fused) {
        auto kfn = &k_set_phase_cells;
        WFM_LAUNCH(kfn, dim3(elementwise_grid(h->ncells)), dim3(256), 0, h->stream, (double*)h->phi.p, (double*)h->s_phi.p,
                   (const double*)h->Z.p, (const uint8_t*)h->mask.p, h->alpha, n, off, h->npix(),
                   (const int*)h->cell_list.p, (const int*)h->in_list.p, h->ncells);
        WFM_CK_LAUNCH(h, "k_set_phase_cells");
        return invalidate_psf(h);                                              // WFM:1648
    }
    auto kfn = &k_set_phase;
    WFM_LAUNCH(kfn, dim3(elementwise_grid(h->npix())), dim3(256), 0, h->stream, (double*)h->phi.p,
               (const double*)h->Z.p, (const uint8_t*)h->mask.p, h->alpha, n, off, h->npix());
    WFM_CK_LAUNCH(h, "k_set_phase");
    h->phi_clean_off_support = true;          // a full pass: phi is zero wherever the mask is off
    return invalidate(h);                                                      // WFM:1648
}

int wfm_set_modulus(wfm_model* h, const double* beta, int n) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_BCAST(h, wfm_set_modulus(c, beta, n));
    if (h->nbatch > 1) return batch_set_modulus(h, beta, n, 0);
    if (n <= 0 || n > WFM_MAX_COEF || !beta) return h->fail(WFM_ERR_INVALID_ARG, "bad modulus coefficient vector");
    if (h->nzern <= 0) return h->fail(WFM_ERR_STATE, "Zernike basis not set");
    if (n > h->nzern)
        return h->fail(WFM_ERR_INVALID_ARG, "DoubleShapedVector beta does not belong to the modulus space");  // WFM:1592
    WFM_ENTER(h);
    double s = 0.0;
    for (int k = 0; k < n; ++k) { h->beta.v[k] = beta[k]; s += beta[k] * beta[k]; }
    h->nmod = n;
    const double beta_norm = 1.0 / std::sqrt(s);                               // WFM:1597
    auto kfn = &k_set_modulus;
    WFM_LAUNCH(kfn, dim3(elementwise_grid(h->npix())), dim3(256), 0, h->stream, (double*)h->rho.p,
               (const double*)h->Z.p, (const uint8_t*)h->mask.p, h->beta, n, beta_norm, h->npix());
    WFM_CK_LAUNCH(h, "k_set_modulus");
    h->have_rho = true;
    return invalidate(h);                                                      // WFM:1609
}

int wfm_set_defocus(wfm_model* h, const double* defoc, int n) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_BCAST(h, wfm_set_defocus(c, defoc, n));
    if (h->nbatch > 1) return batch_set_defocus(h, defoc, n, 0);
    if (!defoc || (n != 1 && n != 3)) return h->fail(WFM_ERR_INVALID_ARG, "bad defocus  parameters");   // WFM:1530, Q4
    if (!h->have_optics) return h->fail(WFM_ERR_STATE, "optics not set: call wfm_set_optics first");
    WFM_ENTER(h);
    if (n == 3) { h->deltaX = defoc[1]; h->deltaY = defoc[2]; }                // WFM:1518-1520
    h->lambda_ni = defoc[0];                                                   // WFM:1522
    h->ni = h->lambda_ni * h->lambda;                                          // WFM:1523
    h->ndefocus = n;
    auto kfn = &k_compute_defocus;
    WFM_LAUNCH(kfn, dim3(elementwise_grid(h->npix())), dim3(256), 0, h->stream, (double*)h->psi.p,
               (uint8_t*)h->mask.p, (const uint8_t*)h->map.p, h->N, h->dxy, h->lambda_ni, h->deltaX, h->deltaY);
    WFM_CK_LAUNCH(h, "k_compute_defocus");
    return invalidate(h);                                                      // WFM:1533
}

// Batch handles: one table row per model.
int wfm_batch_set_phase(wfm_model* h, const double* alpha, int n) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_BCAST(h, wfm_set_phase(c, alpha, n));
    if (h->nbatch <= 1) return wfm_set_phase(h, alpha, n);
    return batch_set_phase(h, alpha, n, n);
}
int wfm_batch_set_modulus(wfm_model* h, const double* beta, int n) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_BCAST(h, wfm_set_modulus(c, beta, n));
    if (h->nbatch <= 1) return wfm_set_modulus(h, beta, n);
    return batch_set_modulus(h, beta, n, n);
}
int wfm_batch_set_defocus(wfm_model* h, const double* defoc, int n) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_BCAST(h, wfm_set_defocus(c, defoc, n));
    if (h->nbatch <= 1) return wfm_set_defocus(h, defoc, n);
    return batch_set_defocus(h, defoc, n, n);
}

int wfm_set_pupil_arrays(wfm_model* h, const double* rho, const double* phi, const double* psi, const uint8_t* mask) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_BCAST(h, wfm_set_pupil_arrays(c, rho, phi, psi, mask));
    if (h->nbatch > 1) return h->fail(WFM_ERR_UNSUPPORTED, "wfm_set_pupil_arrays: not available on a batch handle");
    WFM_ENTER(h);
    const size_t npix = h->npix();
    WFM_CK(h, cudaStreamSynchronize(h->stream));
    if (h->h_esc.empty()) h->h_esc.assign(npix, 0);
    if (rho) {
        WFM_CK(h, cudaMemcpy(h->rho.p, rho, 8 * npix, cudaMemcpyHostToDevice));
        for (size_t i = 0; i < npix; ++i) if (rho[i] != 0.0) h->h_esc[i] = 1;
        h->have_rho = true;
    }
    if (phi) {
        WFM_CK(h, cudaMemcpy(h->phi.p, phi, 8 * npix, cudaMemcpyHostToDevice));
        h->phi_clean_off_support = false;     // arbitrary values anywhere: the next setPhase() takes the full pass
    }
    if (psi) WFM_CK(h, cudaMemcpy(h->psi.p, psi, 8 * npix, cudaMemcpyHostToDevice));
    if (mask) {
        std::vector<uint8_t> m(npix);
        for (size_t i = 0; i < npix; ++i) { m[i] = mask[i] ? 1 : 0; if (m[i]) h->h_esc[i] = 1; }
        WFM_CK(h, cudaMemcpy(h->mask.p, m.data(), npix, cudaMemcpyHostToDevice));
    }
    h->activity_dirty = true;
    return invalidate(h);
}

int wfm_set_modulus_mode(wfm_model* h, int mode) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_BCAST(h, wfm_set_modulus_mode(c, mode));
    if (mode != WFM_MODULUS_INTENDED && mode != WFM_MODULUS_REFERENCE_LAST_PLANE)
        return h->fail(WFM_ERR_INVALID_ARG, "bad modulus mode");
    h->modulus_mode = mode;
    return WFM_OK;
}

static int ensure_copy_stream(wfm_model* h) {
    if (h->copy_stream) return WFM_OK;
    WFM_CK(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    WFM_CK(h, cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
    WFM_CK(h, cudaEventCreateWithFlags(&h->ev_copied, cudaEventDisableTiming));
    return WFM_OK;
}

#define WFM_STAGED_MIN_BYTES ((size_t)8 << 20)
static bool use_staging(const void* host, size_t bytes) {
    size_t min_bytes = WFM_STAGED_MIN_BYTES;
    if (const char* e = getenv("WFM_STAGED_MIN_BYTES")) min_bytes = (size_t)atoll(e);
    return bytes >= min_bytes && host_is_pageable(host);
}

static int copy_out(wfm_model* h, void* out, const void* dev, size_t bytes) {
    if (!out) return h->fail(WFM_ERR_INVALID_ARG, "output pointer is NULL");
    WFM_ENTER(h);
    HostStage* st = use_staging(out, bytes) ? host_stage(h) : nullptr;
    if (st) {                                     // large pageable destination: staged, multi-threaded
        int rc = ensure_copy_stream(h); if (rc) return rc;
        WFM_CK(h, cudaEventRecord(h->ev_ready, h->stream));
        WFM_CK(h, cudaStreamWaitEvent(h->copy_stream, h->ev_ready, 0));
        rc = staged_d2h(h, st, out, dev, bytes, h->copy_stream, [](size_t) { return (int)WFM_OK; });
        if (rc) return rc;
        WFM_CK(h, cudaStreamSynchronize(h->stream));
        return check_pipeline(h);
    }
    WFM_CK(h, cudaMemcpyAsync(out, dev, bytes, cudaMemcpyDeviceToHost, h->stream));
    WFM_CK(h, cudaStreamSynchronize(h->stream));
    return check_pipeline(h);
}

// (batch handles: nbatch arrays, model after model)
int wfm_get_rho(wfm_model* h, double* out) { if (h && h->multi()) h = h->parts[0]; return h ? copy_out(h, out, h->rho.p, 8 * (size_t)h->npix() * h->nbatch) : WFM_ERR_INVALID_ARG; }
int wfm_get_phi(wfm_model* h, double* out) { if (h && h->multi()) h = h->parts[0]; return h ? copy_out(h, out, h->phi.p, 8 * (size_t)h->npix() * h->nbatch) : WFM_ERR_INVALID_ARG; }
int wfm_get_psi(wfm_model* h, double* out) { if (h && h->multi()) h = h->parts[0]; return h ? copy_out(h, out, h->psi.p, 8 * (size_t)h->npix() * h->nbatch) : WFM_ERR_INVALID_ARG; }
int wfm_get_mask(wfm_model* h, uint8_t* out) { if (h && h->multi()) h = h->parts[0]; return h ? copy_out(h, out, h->mask.p, (size_t)h->npix() * h->nbatch) : WFM_ERR_INVALID_ARG; }

int wfm_compute_psf(wfm_model* h) {
    if (!h) return WFM_ERR_INVALID_ARG;
    if (h->multi()) return wfm_multi::compute_psf(h);
    WFM_ENTER(h);
    return compute_psf_impl(h);
}
int wfm_invalidate(wfm_model* h) {
    if (!h) return WFM_ERR_INVALID_ARG;
    for (wfm_model* c : h->parts) invalidate_psf(c);
    return invalidate_psf(h);
}
int wfm_psf_state(const wfm_model* h) {
    if (!h) return 0;
    if (!h->multi()) return h->pstate;
    for (const wfm_model* c : h->parts) if (c->pstate < 1) return 0;
    return 1;
}

int wfm_get_psf(wfm_model* h, void* out) {
    if (!h) return WFM_ERR_INVALID_ARG;
    if (h->multi()) return wfm_multi::get_stack(h, out, false, false);
    WFM_ENTER(h);
    if (!out) return h->fail(WFM_ERR_INVALID_ARG, "output pointer is NULL");
    const size_t plane = (size_t)h->npix() * h->esz(), bytes = plane * h->nzl;
    const int cp = (h->pstate > 0) ? 0 : host_chunk_planes(h);
    HostStage* st = (cp > 0 && use_staging(out, bytes)) ? host_stage(h) : nullptr;
    if (st) {
        // dirty PSF into a large pageable array: computePsf in plane windows (all queued at once), the staged read-back
        // of a piece waits for the window that holds its last byte
        int rc = ensure_copy_stream(h); if (rc) return rc;
        const int nwin = (h->nzl + cp - 1) / cp;
        rc = ensure_chunk_events(h, h->ev_chunk, nwin); if (rc) return rc;
        rc = compute_psf_windows(h, cp, [&](int p0, int) {
            WFM_CK(h, cudaEventRecord(h->ev_chunk[p0 / cp], h->stream));
            return (int)WFM_OK;
        });
        if (rc) return rc;
        const size_t wbytes = plane * cp;
        cudaStream_t cs = h->copy_stream;
        const std::vector<cudaEvent_t>& evs = h->ev_chunk;
        rc = staged_d2h(h, st, out, h->psf.p, bytes, cs, [&evs, cs, wbytes](size_t b) {
            return cudaStreamWaitEvent(cs, evs[(b - 1) / wbytes], 0) == cudaSuccess ? (int)WFM_OK : (int)WFM_ERR_CUDA;
        });
        if (rc) return rc;
        WFM_CK(h, cudaStreamSynchronize(h->stream));
        return check_pipeline(h);
    }
    int rc = compute_psf_impl(h); if (rc) return rc;                           // WFM:1800-1802
    return copy_out(h, out, h->psf.p, bytes);
}

// getPsf() whose device->host copy runs on the handle's second stream: returns once the copy is queued.
// `out` must stay valid (and should be pinned, wfm_host_alloc) until wfm_wait_transfers() returns.  The next
// computePsf() on this handle is ordered after the copy, so the slab is never overwritten while it is read.
int wfm_get_psf_async(wfm_model* h, void* out) {
    if (!h) return WFM_ERR_INVALID_ARG;
    if (h->multi()) return wfm_multi::get_stack(h, out, false, true);
    WFM_ENTER(h);
    if (!out) return h->fail(WFM_ERR_INVALID_ARG, "output pointer is NULL");
    { int rc0 = ensure_copy_stream(h); if (rc0) return rc0; }
    const int cp = (h->pstate > 0) ? 0 : host_chunk_planes(h);
    if (cp > 0) {
        // dirty PSF: computePsf in plane windows, every window's read-back queued right behind it on the second stream
        const size_t plane = (size_t)h->npix() * h->esz();
        int rc = ensure_chunk_events(h, h->ev_chunk, (h->nzl + cp - 1) / cp); if (rc) return rc;
        rc = compute_psf_windows(h, cp, [&](int p0, int np) {
            cudaEvent_t ev = h->ev_chunk[p0 / cp];
            WFM_CK(h, cudaEventRecord(ev, h->stream));
            WFM_CK(h, cudaStreamWaitEvent(h->copy_stream, ev, 0));
            WFM_CK(h, cudaMemcpyAsync((char*)out + plane * p0, (const char*)h->psf.p + plane * p0, plane * np,
                                      cudaMemcpyDeviceToHost, h->copy_stream));
            return (int)WFM_OK;
        });
        if (rc) return rc;
    } else {
        int rc = compute_psf_impl(h); if (rc) return rc;                       // WFM:1800-1802
        WFM_CK(h, cudaEventRecord(h->ev_ready, h->stream));
        WFM_CK(h, cudaStreamWaitEvent(h->copy_stream, h->ev_ready, 0));
        WFM_CK(h, cudaMemcpyAsync(out, h->psf.p, (size_t)h->npix() * h->nzl * h->esz(), cudaMemcpyDeviceToHost, h->copy_stream));
    }
    WFM_CK(h, cudaEventRecord(h->ev_copied, h->copy_stream));
    h->copy_pending = true;
    return WFM_OK;
}

int wfm_wait_transfers(wfm_model* h) {
    if (!h) return WFM_ERR_INVALID_ARG;
    if (h->multi()) return wfm_multi::wait_transfers(h);
    WFM_ENTER(h);
    if (h->copy_stream) WFM_CK(h, cudaStreamSynchronize(h->copy_stream));
    h->copy_pending = false;
    return check_pipeline(h);
}

// ArrayUtils.roll(pupil.getPsf()) (BlindDeconvJob.java:100) -- "next" row f4: the centred PSF, shifted on the device.
int wfm_roll_psf_dev(wfm_model* h, void* out_dev) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_NO(h, "the rolled PSF (the z roll crosses devices)");
    WFM_ENTER(h);
    if (!out_dev) return h->fail(WFM_ERR_INVALID_ARG, "output pointer is NULL");
    if (h->z0 != 0 || h->nzl != h->nz_global)
        return h->fail(WFM_ERR_UNSUPPORTED, "the rolled PSF needs the whole stack of ONE model on the handle (z roll crosses slabs)");
    int rc = compute_psf_impl(h); if (rc) return rc;
    const size_t vox = (size_t)h->npix() * h->nzl;
    const unsigned grid = (unsigned)((vox + 255) / 256);
    if (h->precision == WFM_F64) {
        auto kfn = &k_roll3<double>;
        WFM_LAUNCH(kfn, dim3(grid), dim3(256), 0, h->stream, (double*)out_dev, (const double*)h->psf.p, h->N, h->N, h->nzl);
    } else {
        auto kfn = &k_roll3<float>;
        WFM_LAUNCH(kfn, dim3(grid), dim3(256), 0, h->stream, (float*)out_dev, (const float*)h->psf.p, h->N, h->N, h->nzl);
    }
    WFM_CK_LAUNCH(h, "k_roll3");
    return WFM_OK;
}

int wfm_get_psf_rolled(wfm_model* h, void* out) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_NO(h, "the rolled PSF (the z roll crosses devices)");
    if (!out) return h->fail(WFM_ERR_INVALID_ARG, "output pointer is NULL");
    WFM_ENTER(h);
    const size_t bytes = (size_t)h->npix() * h->nzl * h->esz();
    WFM_CK(h, h->qdev.ensure(bytes));                    // reuse the q staging buffer
    int rc = wfm_roll_psf_dev(h, h->qdev.p); if (rc) return rc;
    return copy_out(h, out, h->qdev.p, bytes);
}

int wfm_get_cpx_psf(wfm_model* h, void* out) {
    if (!h) return WFM_ERR_INVALID_ARG;
    if (h->multi()) return wfm_multi::get_stack(h, out, true, false);
    WFM_ENTER(h);
    int rc = compute_psf_impl(h); if (rc) return rc;                           // WFM:1857-1859
    return copy_out(h, out, h->cpx.p, (size_t)h->npix() * h->nzl * 2 * h->esz());
}

int wfm_device_psf(wfm_model* h, void** p) {
    if (!h || !p) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_NO(h, "wfm_device_psf (use wfm_multi_part and ask the child)");
    WFM_ENTER(h);
    int rc = compute_psf_impl(h); if (rc) return rc;
    *p = h->psf.p; return WFM_OK;
}
int wfm_device_cpx_psf(wfm_model* h, void** p) {
    if (!h || !p) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_NO(h, "wfm_device_cpx_psf (use wfm_multi_part and ask the child)");
    WFM_ENTER(h);
    int rc = compute_psf_impl(h); if (rc) return rc;
    *p = h->cpx.p; return WFM_OK;
}

int wfm_grad_length(const wfm_model* h) { return h ? h->glen() : 0; }

int wfm_apply_jacobian_dev(wfm_model* h, unsigned kinds, const void* q_dev, double* grad_dev) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_NO(h, "wfm_apply_jacobian_dev (use wfm_multi_apply_jacobian_dev)");
    WFM_ENTER(h);
    if (!q_dev || !grad_dev) return h->fail(WFM_ERR_INVALID_ARG, "q_dev / grad_dev is NULL");
    int rc = jacobian_preconditions(h, kinds); if (rc) return rc;
    rc = compute_psf_impl(h); if (rc) return rc;                               // quirk Q5: recompute if dirty
    return (h->precision == WFM_F64) ? dispatch_jac<double>(h, kinds, q_dev, grad_dev)
                                     : dispatch_jac<float>(h, kinds, q_dev, grad_dev);
}


// One device, q in host memory: H2D of q, Jacobian, D2H of the K-vector ([nbatch][glen] doubles into g_out).
// Large slabs move in plane chunks on a stream of their own (host_chunk_planes): the copy starts at once -- beside a
// PSF that is still being computed or read back -- and the adjoint pipeline of chunk c runs under the copy of chunk c+1.
static int apply_host_single(wfm_model* h, unsigned kinds, const void* q_host, double* g_out) {
    WFM_ENTER(h);
    int rc = jacobian_preconditions(h, kinds); if (rc) return rc;
    const size_t plane = (size_t)h->npix() * h->esz(), bytes = plane * h->nzl;
    WFM_CK(h, h->qdev.ensure(bytes));
    WFM_CK(h, h->grad.ensure(8 * (size_t)h->glen() * h->nbatch));
    const int cp = host_chunk_planes(h);
    HostStage* st = use_staging(q_host, bytes) ? host_stage(h) : nullptr;
    if (st) {
        // large pageable q: staged through pinned slots by the host threads; the adjoint window of a plane chunk is
        // launched as soon as the pieces that hold it have been queued
        if (!h->in_stream) WFM_CK(h, cudaStreamCreateWithFlags(&h->in_stream, cudaStreamNonBlocking));
        const int wp = cp > 0 ? cp : h->nzl;                                   // planes per window (one window: the slab)
        const int nwin = (h->nzl + wp - 1) / wp;
        rc = ensure_chunk_events(h, h->ev_in, nwin); if (rc) return rc;
        rc = compute_psf_impl(h); if (rc) return rc;                           // quirk Q5 (queued; the staging overlaps it)
        int next = 0;
        const size_t piece = st->piece();
        rc = staged_h2d(h, st, h->qdev.p, q_host, bytes, h->in_stream, [&](int k) {
            const size_t have = std::min(bytes, (size_t)(k + 1) * piece);
            while (next < nwin && plane * (size_t)std::min(h->nzl, (next + 1) * wp) <= have) {
                const int p0 = next * wp, np = std::min(wp, h->nzl - p0);
                WFM_CK(h, cudaEventRecord(h->ev_in[next], h->in_stream));
                WFM_CK(h, cudaStreamWaitEvent(h->stream, h->ev_in[next], 0));
                if (nwin > 1) { h->win0 = p0; h->winN = np; }
                const int r = (h->precision == WFM_F64) ? dispatch_jac<double>(h, kinds, h->qdev.p, (double*)h->grad.p)
                                                        : dispatch_jac<float>(h, kinds, h->qdev.p, (double*)h->grad.p);
                h->win0 = 0; h->winN = 0;
                if (r) return r;
                ++next;
            }
            return (int)WFM_OK;
        });
        if (rc) { cudaStreamSynchronize(h->in_stream); cudaStreamSynchronize(h->stream); return rc; }
    } else if (cp <= 0) {
        WFM_CK(h, cudaMemcpyAsync(h->qdev.p, q_host, bytes, cudaMemcpyHostToDevice, h->stream));
        rc = wfm_apply_jacobian_dev(h, kinds, h->qdev.p, (double*)h->grad.p); if (rc) return rc;
    } else {
        if (!h->in_stream) WFM_CK(h, cudaStreamCreateWithFlags(&h->in_stream, cudaStreamNonBlocking));
        const int nch = (h->nzl + cp - 1) / cp;
        rc = ensure_chunk_events(h, h->ev_in, nch); if (rc) return rc;
        // (qdev is idle here: every call that reads it drains the handle's stream before it returns)
        for (int c = 0; c < nch; ++c) {
            const int p0 = c * cp, np = std::min(cp, h->nzl - p0);
            WFM_CK(h, cudaMemcpyAsync((char*)h->qdev.p + plane * p0, (const char*)q_host + plane * p0, plane * np,
                                      cudaMemcpyHostToDevice, h->in_stream));
            WFM_CK(h, cudaEventRecord(h->ev_in[c], h->in_stream));
        }
        rc = compute_psf_impl(h);                                              // quirk Q5: recompute if dirty
        for (int c = 0; c < nch && !rc; ++c) {
            const int p0 = c * cp, np = std::min(cp, h->nzl - p0);
            WFM_CK(h, cudaStreamWaitEvent(h->stream, h->ev_in[c], 0));
            h->win0 = p0; h->winN = np;
            rc = (h->precision == WFM_F64) ? dispatch_jac<double>(h, kinds, h->qdev.p, (double*)h->grad.p)
                                           : dispatch_jac<float>(h, kinds, h->qdev.p, (double*)h->grad.p);
            h->win0 = 0; h->winN = 0;
        }
        if (rc) { cudaStreamSynchronize(h->in_stream); return rc; }
    }
    WFM_CK(h, cudaMemcpyAsync(g_out, h->grad.p, 8 * (size_t)h->glen() * h->nbatch, cudaMemcpyDeviceToHost, h->stream));
    WFM_CK(h, cudaStreamSynchronize(h->stream));
    return check_pipeline(h);
}

static int apply_host(wfm_model* h, unsigned kinds, const void* q_host, std::vector<double>& g) {
    if (!q_host) return h->fail(WFM_ERR_INVALID_ARG, "q is NULL");
    if (h->multi()) return wfm_multi::apply_host(h, kinds, q_host, g);
    g.resize((size_t)h->glen() * h->nbatch);
    return apply_host_single(h, kinds, q_host, g.data());
}

#define WFM_NO_BATCH(h) do { if ((h)->nbatch > 1) return (h)->fail(WFM_ERR_UNSUPPORTED, "batch handle: use wfm_batch_apply_jacobian"); } while (0)

int wfm_apply_j_phase(wfm_model* h, const void* q, double* out, int n) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_NO_BATCH(h);
    if (!out || n != h->nphase || n <= 0) return h->fail(WFM_ERR_INVALID_ARG, "output length must equal nPhase");
    std::vector<double> g;
    int rc = apply_host(h, WFM_J_PHASE, q, g); if (rc) return rc;
    memcpy(out, g.data() + 3, 8 * (size_t)n);
    return WFM_OK;
}

int wfm_apply_j_defocus(wfm_model* h, const void* q, double* out, int n) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_NO_BATCH(h);
    if (!out || (n != 1 && n != 3)) return h->fail(WFM_ERR_INVALID_ARG, "defocus gradient has 1 or 3 elements");   // Q4
    std::vector<double> g;
    int rc = apply_host(h, WFM_J_DEFOCUS, q, g); if (rc) return rc;
    memcpy(out, g.data(), 8 * (size_t)n);                                      // WFM:1352-1359
    return WFM_OK;
}

int wfm_apply_j_modulus(wfm_model* h, const void* q, double* out, int n) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_NO_BATCH(h);
    if (!out || n != h->nmod || n <= 0) return h->fail(WFM_ERR_INVALID_ARG, "output length must equal nModulus");
    std::vector<double> g;
    int rc = apply_host(h, WFM_J_MODULUS, q, g); if (rc) return rc;
    memcpy(out, g.data() + 3 + h->nphase, 8 * (size_t)n);
    return WFM_OK;
}

int wfm_apply_jacobian(wfm_model* h, int param, const void* q, double* out, int n) {
    if (!h) return WFM_ERR_INVALID_ARG;
    switch (param) {                                                           // WFM:399-409
        case WFM_DEFOCUS: return wfm_apply_j_defocus(h, q, out, n);
        case WFM_PHASE: return wfm_apply_j_phase(h, q, out, n);
        case WFM_MODULUS: return wfm_apply_j_modulus(h, q, out, n);
        default: return h->fail(WFM_ERR_INVALID_ARG, "DoubleShapedVector grad does not belong to any space");
    }
}

// Batch handles: the Jacobians of every model in one pass.  q_host = [nbatch][Nz][Ny][Nx] like the PSF;
// out = [nbatch][3 + nPhase + nModulus] = [defocus(3) | phase | modulus] per model (zeros for kinds not selected).
int wfm_batch_apply_jacobian(wfm_model* h, unsigned kinds, const void* q, double* out) {
    if (!h) return WFM_ERR_INVALID_ARG;
    if (!out) return h->fail(WFM_ERR_INVALID_ARG, "output pointer is NULL");
    std::vector<double> g;
    int rc = apply_host(h, kinds, q, g); if (rc) return rc;
    memcpy(out, g.data(), 8 * g.size());
    return WFM_OK;
}

int wfm_apply_j_all(wfm_model* h, const void* q, double* d3, double* ph, double* mo) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_NO_BATCH(h);
    unsigned kinds = 0;
    if (d3) kinds |= WFM_J_DEFOCUS;
    if (ph) kinds |= WFM_J_PHASE;
    if (mo) kinds |= WFM_J_MODULUS;
    std::vector<double> g;
    int rc = apply_host(h, kinds, q, g); if (rc) return rc;
    if (d3) memcpy(d3, g.data(), 24);
    if (ph) memcpy(ph, g.data() + 3, 8 * (size_t)h->nphase);
    if (mo) memcpy(mo, g.data() + 3 + h->nphase, 8 * (size_t)h->nmod);
    return WFM_OK;
}

int wfm_fill_uniform(wfm_model* h, void* dev, int precision, uint64_t seed, uint64_t first, uint64_t count) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_NO(h, "wfm_fill_uniform (use wfm_multi_part and fill on the child's device)");
    if (!dev) return h->fail(WFM_ERR_INVALID_ARG, "dev_ptr is NULL");
    WFM_ENTER(h);
    const unsigned grid = (unsigned)((count + 255) / 256);
    if (precision == WFM_F64) {
        auto kfn = &k_fill_uniform<double>;
        WFM_LAUNCH(kfn, dim3(grid), dim3(256), 0, h->stream, (double*)dev, seed, first, count);
    } else {
        auto kfn = &k_fill_uniform<float>;
        WFM_LAUNCH(kfn, dim3(grid), dim3(256), 0, h->stream, (float*)dev, seed, first, count);
    }
    WFM_CK_LAUNCH(h, "k_fill_uniform");
    return WFM_OK;
}

int wfm_host_alloc(void** out, size_t bytes) {
    if (!out) return WFM_ERR_INVALID_ARG;
    return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? WFM_OK : WFM_ERR_NOMEM;
}
int wfm_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? WFM_OK : WFM_ERR_CUDA; }

int wfm_get_info(const wfm_model* h, int* nx, int* ny, int* nzg, int* z0, int* nzl, int* prec, int* nzern,
                 int* nphase, int* nmod) {
    if (!h) return WFM_ERR_INVALID_ARG;
    if (nx) *nx = h->N;
    if (ny) *ny = h->N;
    if (nzg) *nzg = h->nz_global;
    if (z0) *z0 = h->z0;
    if (nzl) *nzl = h->nzl;
    if (prec) *prec = h->precision;
    if (nzern) *nzern = h->nzern;
    if (nphase) *nphase = h->nphase;
    if (nmod) *nmod = h->nmod;
    return WFM_OK;
}

int wfm_active_extent(const wfm_model* h, int* nax, int* nay) {
    if (!h) return WFM_ERR_INVALID_ARG;
    if (h->multi()) return wfm_active_extent(h->parts[0], nax, nay);
    DeviceScope dev_scope__(h->device);
    int rc = rebuild_activity(const_cast<wfm_model*>(h)); if (rc) return rc;
    if (nax) *nax = h->nax;
    if (nay) *nay = h->nay;
    return WFM_OK;
}

int wfm_set_profiling(wfm_model* h, int on) {
    if (!h) return WFM_ERR_INVALID_ARG;
    WFM_MULTI_BCAST(h, wfm_set_profiling(c, on));
    WFM_ENTER(h);
    WFM_CK(h, cudaStreamSynchronize(h->stream));
    drain_spans(h);
    h->profiling = on != 0;
    for (int i = 0; i < WFM_KERNEL_IDS; ++i) { h->k_ms[i] = 0.0; h->k_count[i] = 0; }
    return WFM_OK;
}

int wfm_get_kernel_times(wfm_model* h, double* ms, uint64_t* counts, int n) {
    if (!h || !ms || !counts || n < WFM_KERNEL_IDS) return WFM_ERR_INVALID_ARG;
    if (h->multi()) return wfm_multi::kernel_times(h, ms, counts);
    WFM_ENTER(h);
    WFM_CK(h, cudaStreamSynchronize(h->stream));
    drain_spans(h);
    for (int i = 0; i < WFM_KERNEL_IDS; ++i) { ms[i] = h->k_ms[i]; counts[i] = h->k_count[i]; }
    return WFM_OK;
}

uint64_t wfm_launch_count(void) {
#ifdef WFM_EMU
    return emu::launch_counter().load();
#else
    return wfm::launch_counter().load();
#endif
}

const char* wfm_version(void) {
#ifdef WFM_EMU
    return "microtipi_b200 0.1 (CPU emulation build: tests only)";
#else
    return "microtipi_b200 0.1 (sm_100a)";
#endif
}

}  // extern "C"

#include "wfm_conv_api.inl"
#include "wfm_multi.inl"
#include "wfm_conv_multi.inl"
