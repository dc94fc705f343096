// Shared helpers for the DIGAT sm_100a kernels: error reporting across the C ABI, launch checks, small device utils.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/digat_sm100.h"

namespace digat {

static thread_local char g_last_error[512] = "";

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(DIGAT_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return DIGAT_OK;
}

#define DIGAT_CUDA(call)                                                                          \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) return ::digat::fail(DIGAT_E_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

#define DIGAT_REQUIRE(cond, ...)                                           \
    do {                                                                   \
        if (!(cond)) return ::digat::fail(DIGAT_E_INVALID, __VA_ARGS__);   \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Device properties are queried once per device and cached (the only mutable global state besides TMA descriptors).
struct DeviceInfo {
    int sm_count = 0;
    int cc_major = 0;
    int max_smem_optin = 0;
    bool valid = false;
};

inline const DeviceInfo* device_info() {
    static DeviceInfo infos[16];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    DeviceInfo& d = infos[dev];
    if (!d.valid) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return nullptr;
        d.sm_count = p.multiProcessorCount;
        d.cc_major = p.major;
        d.max_smem_optin = (int)p.sharedMemPerBlockOptin;
        d.valid = true;   // benign race: every thread writes the same values
    }
    return &d;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit global accesses (data touched once: keep it out of L1)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

constexpr float kNegFill = -1e9f;     // masked_fill value of the reference (finite on purpose)
constexpr float kLeakySlope = 0.2f;

}  // namespace digat
