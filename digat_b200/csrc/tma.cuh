// mbarrier + TMA (cp.async.bulk.tensor) helpers shared by the tcgen05 GEMM and the fused graph-layer kernel, and the
// host-side tensor-map encoder (cuTensorMapEncodeTiled through the runtime's driver entry point: no -lcuda needed).
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <unordered_map>
#include "common.cuh"

namespace digat {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
constexpr uint32_t kMbarSuspendHintNs = 20000u;
// Bounded spin: a barrier that never completes traps (CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; ; ++spin) {
        // suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint expires) instead of
        // re-issuing the poll -- the spin loops of waiting warps were 15 % of all issued instructions of the layer kernel
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity), "r"(kMbarSuspendHintNs) : "memory");
        if (done) return;
        if (spin > (1u << 23)) {
            printf("digat: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
            __trap();
        }
    }
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

#ifdef DIGAT_FAKE_TMA_ENCODER
inline int& fake_encode_calls() { static int n = 0; return n; }
#endif
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                             CUtensorMapFloatOOBfill);

#ifdef DIGAT_FAKE_TMA_ENCODER   // host-only test harness of the descriptor cache (tests/cache_harness.cu): no driver needed
inline CUresult fake_tensor_map_encode(CUtensorMap* m, CUtensorMapDataType, cuuint32_t, void* base, const cuuint64_t* gdim,
                                       const cuuint64_t* gstride, const cuuint32_t* box, const cuuint32_t*,
                                       CUtensorMapInterleave, CUtensorMapSwizzle sw, CUtensorMapL2promotion,
                                       CUtensorMapFloatOOBfill) {
    uint64_t* w = reinterpret_cast<uint64_t*>(m);
    for (int i = 0; i < 16; ++i) w[i] = 0;
    w[0] = reinterpret_cast<uintptr_t>(base); w[1] = gdim[0]; w[2] = gdim[1]; w[3] = gstride[0]; w[4] = box[0]; w[5] = box[1]; w[6] = sw;
    ++fake_encode_calls();
    return CUDA_SUCCESS;
}
#endif

inline PFN_tensorMapEncodeTiled tensor_map_encoder() {
#ifdef DIGAT_FAKE_TMA_ENCODER
    return &fake_tensor_map_encode;
#endif
    static PFN_tensorMapEncodeTiled fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tensorMapEncodeTiled>(p);
    }
    return fn;
}

// Descriptor cache.  A CUtensorMap is a pure function of (base address, shape, pitch, box, swizzle, element type), so an
// encoded descriptor can be reused for as long as the process lives, whatever happens to the memory behind it (a buffer
// freed and reallocated at the same address with the same shape needs the very same descriptor).  PyTorch's caching
// allocator hands the same blocks to the same call sites step after step, so in steady state no launch encodes anything
// (the driver call costs ~1 us per descriptor, three to four per GEMM launch).  Mutex-guarded; bounded (cleared when full).
struct TensorMapKey {
    const void* base;
    int64_t rows, cols, ld;
    int box_rows, box_cols, swizzle, dtype;
    bool operator==(const TensorMapKey& o) const {
        return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
               box_cols == o.box_cols && swizzle == o.swizzle && dtype == o.dtype;
    }
};
struct TensorMapKeyHash {
    size_t operator()(const TensorMapKey& k) const {
        uint64_t h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
        auto mix = [&](uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
        mix((uint64_t)k.rows); mix((uint64_t)k.cols); mix((uint64_t)k.ld);
        mix(((uint64_t)k.box_rows << 32) | (uint32_t)k.box_cols); mix(((uint64_t)k.swizzle << 8) | (uint32_t)k.dtype);
        return (size_t)h;
    }
};

inline int make_tensor_map_2d_any(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                                  int box_cols, CUtensorMapSwizzle swizzle, CUtensorMapDataType dtype, int elem_bytes) {
    static const bool no_cache = getenv("DIGAT_NO_DESC_CACHE") != nullptr;          // debugging aid
    struct Blob { unsigned char bytes[sizeof(CUtensorMap)]; };   // plain bytes: CUtensorMap is alignas(64), map nodes need not be
    static std::mutex mu;
    static std::unordered_map<TensorMapKey, Blob, TensorMapKeyHash> cache;
    const TensorMapKey key{base, rows, cols, ld, box_rows, box_cols, (int)swizzle, (int)dtype};
    if (!no_cache) {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            memcpy(map, it->second.bytes, sizeof(CUtensorMap));
            return DIGAT_OK;
        }
    }
    PFN_tensorMapEncodeTiled enc = tensor_map_encoder();
    if (!enc) return fail(DIGAT_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * (cuuint64_t)elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dtype, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DIGAT_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    if (no_cache) return DIGAT_OK;
    Blob blob;
    memcpy(blob.bytes, map, sizeof(CUtensorMap));
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() >= 4096) cache.clear();
    cache.emplace(key, blob);
    return DIGAT_OK;
}

// 2-D fp32 row-major [rows, cols] with row pitch ld elements; box = [box_rows, box_cols]; out-of-bounds reads give 0.
inline int make_tensor_map_2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                              int box_cols, CUtensorMapSwizzle swizzle) {
    return make_tensor_map_2d_any(map, base, rows, cols, ld, box_rows, box_cols, swizzle, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4);
}

// [rows, cols] bf16 (2-byte elements), row pitch ld elements
inline int make_tensor_map_2d_bf16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                                   int box_cols, CUtensorMapSwizzle swizzle) {
    return make_tensor_map_2d_any(map, base, rows, cols, ld, box_rows, box_cols, swizzle, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel instantiation, device) instead of once per launch.
template <typename Kernel>
inline int ensure_dynamic_smem(Kernel kernel, size_t bytes) {
    static std::mutex mu;
    static std::unordered_map<uint64_t, size_t> done;          // (kernel address ^ device) -> largest size set
    int dev = 0;
    DIGAT_CUDA(cudaGetDevice(&dev));
    const uint64_t key = (uint64_t)reinterpret_cast<uintptr_t>(reinterpret_cast<const void*>(kernel)) * 31u + (uint64_t)dev;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = done.find(key);
        if (it != done.end() && it->second >= bytes) return DIGAT_OK;
    }
    DIGAT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    std::lock_guard<std::mutex> lock(mu);
    done[key] = bytes;
    return DIGAT_OK;
}

}  // namespace digat
