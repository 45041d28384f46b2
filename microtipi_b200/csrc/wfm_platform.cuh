// wfm_platform.cuh -- the two macros through which kernels touch the launch machinery.
//
// Product build (nvcc, sm_100a): real dynamic shared memory and <<< >>> launches.
// Test build (-DWFM_EMU with tests/emu/cuda_emu.h force-included, g++): the CPU emulator
// supplies both macros so that the identical kernel sources can be checked without a GPU.
#pragma once

#ifndef WFM_EMU
#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>

namespace wfm {
inline std::atomic<unsigned long long>& launch_counter() {
    static std::atomic<unsigned long long> c{0};
    return c;
}
}  // namespace wfm

#define WFM_DYN_SMEM(T, name)                                              \
    extern __shared__ __align__(16) unsigned char wfm_dyn_smem_raw[];      \
    T* name = reinterpret_cast<T*>(wfm_dyn_smem_raw)

#define WFM_SPIN_PAUSE() __nanosleep(40)
// wall clock of the device in nanoseconds (%globaltimer): the dependency waits time out on TIME, not on a spin
// count, so that a slow neighbour (co-tenancy, a debugger) cannot turn a long wait into a false error
__device__ __forceinline__ unsigned long long wfm_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// TMA bulk prefetch of a contiguous global range into L2 (no registers, no shared memory, one
// instruction per row): bytes must be a multiple of 16, the address 16-byte aligned.
__device__ __forceinline__ void wfm_prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Drop one 128-byte line from L2 WITHOUT writing it back (its contents become undefined): for ring data that has been
// consumed and will be overwritten before it is read again -- otherwise every dead ring line is written to DRAM when it
// is finally evicted.  p must be 128-byte aligned.  (PTX: the discard is performed as a write of an unspecified value, so
// the barrier + fence + atomic that publishes the ring slot orders it before the next tenant's stores.)
__device__ __forceinline__ void wfm_discard_l2(const void* p) {
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}

// ---- mbarrier + 1-D bulk-async copy (TMA engine; SASS: UBLKCP + SYNCS) ---------------------------------------
// One elected thread posts the expected byte count on the shared-memory barrier and issues the copy; the data moves
// global -> shared without passing through registers or the LSU; every consumer waits on the barrier's phase.
#include <stdint.h>
__device__ __forceinline__ void wfm_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void wfm_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// bytes: multiple of 16; both addresses 16-byte aligned
__device__ __forceinline__ void wfm_bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
// the same in two steps, for several copies completing on one barrier phase
__device__ __forceinline__ void wfm_mbar_expect(uint64_t* bar, unsigned bytes) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wfm_bulk_copy(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void wfm_mbar_complete_emu(uint64_t*) {}   // (the hardware completes the phase by byte count)
// wait until phase number `phase` (0, 1, 2, ... since the init) of the barrier has completed
__device__ __forceinline__ void wfm_mbar_wait(uint64_t* bar, unsigned phase) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar), parity = phase & 1u;
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    } while (!ok);
}

// Programmatic dependent launch: a kernel launched with WFM_LAUNCH_PDL may become resident while its
// predecessor on the stream is still draining; it must execute wfm_grid_dep_wait() before its first access
// to anything the predecessor reads or writes (everything before that point -- twiddle tables into shared
// memory, index set-up -- overlaps the predecessor's tail and the launch latency).
__device__ __forceinline__ void wfm_grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the next PDL kernel on the stream be scheduled as soon as this grid leaves room for it (it still blocks in
// wfm_grid_dep_wait() until this grid has completed and its writes are visible).
__device__ __forceinline__ void wfm_grid_dep_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

namespace wfm {
template <class K, class... Args>
inline void launch_pdl(K kfn, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool off = getenv("WFM_NO_PDL") != nullptr;       // experiment / debugging switch
    cfg.attrs = at; cfg.numAttrs = off ? 0 : 1;
    cudaLaunchKernelEx(&cfg, kfn, args...);
}
}  // namespace wfm
#define WFM_LAUNCH_PDL(kfn, grid, block, smem, stream, ...)                \
    do {                                                                   \
        ::wfm::launch_counter()++;                                         \
        ::wfm::launch_pdl(kfn, (grid), (block), (smem), (stream), __VA_ARGS__); \
    } while (0)

#define WFM_LAUNCH(kfn, grid, block, smem, stream, ...)                    \
    do {                                                                   \
        ::wfm::launch_counter()++;                                         \
        kfn<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);           \
    } while (0)
#endif
