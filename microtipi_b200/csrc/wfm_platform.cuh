// wfm_platform.cuh -- the two macros through which kernels touch the launch machinery.
//
// Product build (nvcc, sm_100a): real dynamic shared memory and <<< >>> launches.
// Test build (-DWFM_EMU with tests/emu/cuda_emu.h force-included, g++): the CPU emulator
// supplies both macros so that the identical kernel sources can be checked without a GPU.
#pragma once

#ifndef WFM_EMU
#include <cuda_runtime.h>
#include <atomic>

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

// TMA bulk prefetch of a contiguous global range into L2 (no registers, no shared memory, one
// instruction per row): bytes must be a multiple of 16, the address 16-byte aligned.
__device__ __forceinline__ void wfm_prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}


#define WFM_LAUNCH(kfn, grid, block, smem, stream, ...)                    \
    do {                                                                   \
        ::wfm::launch_counter()++;                                         \
        kfn<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);           \
    } while (0)
#endif
