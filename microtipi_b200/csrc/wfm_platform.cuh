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

// Ring accesses: bypass L1 (the slot is rewritten by other SMs) and ask L2 to keep the lines
// (evict_last), while the streamed inputs/outputs use the .cs (evict-first) operators.
namespace wfm {
__device__ __forceinline__ unsigned long long ring_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double2 ring_load(const double2* p, unsigned long long pol) {
    double2 r;
    asm volatile("ld.global.cg.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float2 ring_load(const float2* p, unsigned long long pol) {
    float2 r;
    asm volatile("ld.global.cg.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(r.x), "=f"(r.y) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void ring_store(double2* p, double2 v, unsigned long long pol) {
    asm volatile("st.global.cg.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void ring_store(float2* p, float2 v, unsigned long long pol) {
    asm volatile("st.global.cg.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
}  // namespace wfm

#define WFM_LAUNCH(kfn, grid, block, smem, stream, ...)                    \
    do {                                                                   \
        ::wfm::launch_counter()++;                                         \
        kfn<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);           \
    } while (0)
#endif
