// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal CUDA-on-CPU emulator so that the *same* kernel sources that nvcc compiles for
// sm_100a (microtipi_b200/csrc/*.cuh, wfm_api.cu) can be compiled with g++ and executed in
// this GPU-less container: index math, shared-memory exchange patterns, barriers, warp
// shuffles and the host-side state machine of the C ABI are then checked against the oracle
// by `pytest -m "not gpu"`.  The emulated library (tests/emu/_build/libwfm_emu.so) is never
// loaded by the product package `microtipi_b200`, which binds libwfm_b200.so only and fails
// loudly when it (or a GPU) is missing.
//
// Model: one CTA = one OS thread; every CUDA thread of the CTA is a ucontext fiber on it.
// __syncthreads() and the warp shuffles yield to a round-robin scheduler.  CTAs of a grid run
// in parallel on a small pool of OS threads.
#pragma once
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <ucontext.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define WFM_EMU 1

// ---- qualifiers -------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static thread_local
#define __constant__ static

// ---- vector types -----------------------------------------------------------------------
struct alignas(8) float2 { float x, y; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };

// ---- runtime stubs ------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; int major, minor; size_t l2CacheSize; char name[64]; };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
// "devices": WFM_EMU_DEVICES of them (default 1); all share the host heap, so peer access and IPC mappings are the
// identity -- enough to exercise the multi-device / exchange HOST logic and the exchange protocol of the kernels
static inline int emu_device_count() { const char* e = getenv("WFM_EMU_DEVICES"); const int n = e ? atoi(e) : 1; return n < 1 ? 1 : n; }
static inline int& emu_current_device() { static thread_local int d = 0; return d; }
static inline cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= emu_device_count()) return cudaErrorInvalidValue; emu_current_device() = d; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = emu_current_device(); return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = emu_device_count(); return cudaSuccess; }
enum { cudaErrorPeerAccessAlreadyEnabled = 704, cudaIpcMemLazyEnablePeerAccess = 1 };
static inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
struct cudaIpcMemHandle_t { char reserved[64]; };
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof(*h)); memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof(*p)); return cudaSuccess; }
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof(*p)); p->multiProcessorCount = 4; p->sharedMemPerBlockOptin = 227 * 1024; p->major = 10; p->minor = 0;
    p->l2CacheSize = 126u << 20; snprintf(p->name, sizeof(p->name), "emulated"); return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)0x1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)0x2; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)0x2; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

// ---- device intrinsics ---------------------------------------------------------------------
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

namespace emu {

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
    uint3_emu tid{0, 0, 0};
};

struct Block {
    std::vector<Fiber> fibers;
    ucontext_t sched;
    int current = -1;
    int nthreads = 0;
    // CTA barrier
    int bar_count = 0;
    unsigned bar_gen = 0;
    // per-warp shuffle state
    struct Warp { int count = 0; unsigned gen = 0; uint64_t slot[32]; };
    std::vector<Warp> warps;
    struct Named { int count = 0; unsigned gen = 0; };
    Named named[16];
    uint3_emu bid{0, 0, 0};
    dim3 bdim, gdim;
    char* smem = nullptr;
    unsigned long progress = 0;
    const std::function<void()>* body = nullptr;
};

inline Block*& cur() { static thread_local Block* b = nullptr; return b; }

inline void yield_fiber() {
    Block* b = cur();
    Fiber& f = b->fibers[b->current];
    swapcontext(&f.ctx, &b->sched);
}

inline void syncthreads() {
    Block* b = cur();
    unsigned g = b->bar_gen;
    if (++b->bar_count == b->nthreads) { b->bar_count = 0; b->bar_gen++; b->progress++; }
    else while (b->bar_gen == g) yield_fiber();
}

// bar.sync id, count: `count` threads rendezvous on barrier `id`
inline void named_barrier(int id, int count) {
    Block* b = cur();
    Block::Named& nb = b->named[id & 15];
    unsigned g = nb.gen;
    if (++nb.count == count) { nb.count = 0; nb.gen++; b->progress++; }
    else while (nb.gen == g) yield_fiber();
}

inline int lane_id() { Block* b = cur(); const Fiber& f = b->fibers[b->current]; (void)f; return b->current & 31; }

// all 32 lanes of the warp (or the tail warp's lanes) rendezvous
inline void warp_sync() {
    Block* b = cur();
    int w = b->current >> 5;
    Block::Warp& W = b->warps[w];
    int lanes = std::min(32, b->nthreads - w * 32);
    unsigned g = W.gen;
    if (++W.count == lanes) { W.count = 0; W.gen++; b->progress++; }
    else while (W.gen == g) yield_fiber();
}

template <class T> inline T shfl_generic(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    Block* b = cur();
    int w = b->current >> 5, l = b->current & 31;
    Block::Warp& W = b->warps[w];
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    W.slot[l] = raw;
    warp_sync();
    int lanes = std::min(32, b->nthreads - w * 32);
    uint64_t got = (src_lane >= 0 && src_lane < lanes) ? W.slot[src_lane] : raw;
    warp_sync();
    T out; memcpy(&out, &got, sizeof(T));
    return out;
}

extern "C" inline void fiber_entry() {
    Block* b = cur();
    (*b->body)();
    b->fibers[b->current].done = true;
    b->progress++;
    swapcontext(&b->fibers[b->current].ctx, &b->sched);
}

static constexpr size_t kStack = 96 * 1024;

inline void run_block(Block& b) {
    cur() = &b;
    int remaining = b.nthreads;
    for (int i = 0; i < b.nthreads; ++i) {
        Fiber& f = b.fibers[i];
        f.done = false;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &b.sched;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    b.bar_count = 0;
    for (auto& w : b.warps) w.count = 0;
    for (auto& n : b.named) n.count = 0;
    while (remaining > 0) {
        unsigned long before = b.progress;
        for (int i = 0; i < b.nthreads; ++i) {
            Fiber& f = b.fibers[i];
            if (f.done) continue;
            b.current = i;
            swapcontext(&b.sched, &f.ctx);
            if (f.done) --remaining;
        }
        if (remaining > 0 && b.progress == before) {
            fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): a barrier was not reached by every thread\n", b.bid.x, b.bid.y, b.bid.z);
            abort();
        }
    }
    cur() = nullptr;
}

inline int pool_size() {
    const char* e = getenv("WFM_EMU_THREADS");
    int n = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(n, 16));
}

inline std::atomic<uint64_t>& launch_counter() { static std::atomic<uint64_t> c{0}; return c; }

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    const long nblocks = (long)grid.x * grid.y * grid.z;
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nblocks == 0 || nthreads == 0) return;
    std::atomic<long> next{0};
    auto worker = [&]() {
        Block b;
        b.nthreads = nthreads;
        b.fibers.resize(nthreads);
        b.warps.resize((nthreads + 31) / 32);
        b.bdim = block; b.gdim = grid;
        b.body = &body;
        char* stacks = (char*)malloc(kStack * (size_t)nthreads);
        b.smem = (char*)aligned_alloc(128, ((smem_bytes + 127) / 128 + 1) * 128);
        for (int i = 0; i < nthreads; ++i) {
            b.fibers[i].stack = stacks + kStack * (size_t)i;
            b.fibers[i].tid = uint3_emu{(unsigned)i % block.x, ((unsigned)i / block.x) % block.y, (unsigned)i / (block.x * block.y)};
        }
        for (;;) {
            long id = next.fetch_add(1);
            if (id >= nblocks) break;
            b.bid = uint3_emu{(unsigned)(id % grid.x), (unsigned)((id / grid.x) % grid.y), (unsigned)(id / ((long)grid.x * grid.y))};
            run_block(b);
        }
        free(stacks);
        free(b.smem);
    };
    int nw = (int)std::min<long>(pool_size(), nblocks);
    if (nw <= 1) { worker(); return; }
    std::vector<std::thread> pool;
    for (int i = 0; i < nw; ++i) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
}

}  // namespace emu

#define threadIdx (emu::cur()->fibers[emu::cur()->current].tid)
#define blockIdx (emu::cur()->bid)
#define blockDim (emu::cur()->bdim)
#define gridDim (emu::cur()->gdim)
#define __syncthreads() emu::syncthreads()
#define __syncwarp(...) emu::warp_sync()
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::shfl_generic(v, (emu::cur()->current & 31) ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { int l = emu::cur()->current & 31; return emu::shfl_generic(v, l + d < 32 ? l + d : l); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl_generic(v, src); }

static inline double atomicAdd(double* addr, double v) {
    uint64_t* p = (uint64_t*)addr; uint64_t old = __atomic_load_n(p, __ATOMIC_RELAXED), nv;
    double o;
    do { memcpy(&o, &old, 8); double n = o + v; memcpy(&nv, &n, 8); }
    while (!__atomic_compare_exchange_n(p, &old, nv, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
    return o;
}
static inline unsigned atomicAdd(unsigned* addr, unsigned v) { return __atomic_fetch_add(addr, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicCAS(unsigned* addr, unsigned cmp, unsigned val) {
    __atomic_compare_exchange_n(addr, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
static inline int atomicAdd(int* addr, int v) { return __atomic_fetch_add(addr, v, __ATOMIC_SEQ_CST); }

// polling back-off of the pipeline kernels: let the producer CTA's OS thread run
#include <chrono>
#define WFM_SPIN_PAUSE() std::this_thread::sleep_for(std::chrono::microseconds(50))
static inline unsigned long long wfm_now_ns() {
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static inline void wfm_prefetch_l2(const void*, unsigned) {}
static inline void wfm_discard_l2(const void*) {}
// mbarrier + bulk-async copy: the copy is done on the spot by the issuing fiber; the barrier word counts completed phases
static inline void wfm_mbar_init(uint64_t* bar, unsigned) { __atomic_store_n(bar, (uint64_t)0, __ATOMIC_SEQ_CST); }
static inline void wfm_mbar_init_fence() {}
static inline void wfm_bulk_load(void* d, const void* s, unsigned n, uint64_t* bar) {
    memcpy(d, s, n);
    __atomic_fetch_add(bar, (uint64_t)1, __ATOMIC_SEQ_CST);
}
static inline void wfm_mbar_expect(uint64_t*, unsigned) {}
// (emulation: the LAST copy of a phase completes it -- callers pass `last`)
static inline void wfm_bulk_copy(void* d, const void* s, unsigned n, uint64_t*) { memcpy(d, s, n); }
static inline void wfm_mbar_complete_emu(uint64_t* bar) { __atomic_fetch_add(bar, (uint64_t)1, __ATOMIC_SEQ_CST); }
static inline void wfm_mbar_wait(uint64_t* bar, unsigned phase) {
    while (__atomic_load_n(bar, __ATOMIC_SEQ_CST) <= (uint64_t)phase) emu::yield_fiber();
}
static inline void wfm_grid_dep_wait() {}
static inline void wfm_grid_dep_trigger() {}

// dynamic shared memory of the running CTA
#define WFM_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(emu::cur()->smem)

// kernel launch through a function pointer
#define WFM_LAUNCH(kfn, grid, block, smem, stream, ...)                                   \
    do { emu::launch_counter()++; (void)(stream);                                          \
         emu::launch((grid), (block), (smem), [&]() { kfn(__VA_ARGS__); }); } while (0)
#define WFM_LAUNCH_PDL WFM_LAUNCH
