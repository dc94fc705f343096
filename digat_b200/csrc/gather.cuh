// Row gathers and small glue kernels (replace index_select at reference util.py:34-36, 65-67, the cat at
// graphEncoders.py:179/191 and the logits dot at model.py:76,89).  All HBM-bound: one warp moves one 4*D-byte row
// with 128-bit loads/stores; indices are int32 exactly as the reference stores them (construct_SAG.py:452).
#pragma once
#include "common.cuh"

namespace digat {

constexpr int kGatherThreads = 256;

// out[r, s, :] = table[ level1 ? level1[idx0[r] * n_inner + s] : idx0[r * n_inner + s] , :]
__global__ void __launch_bounds__(kGatherThreads)
gather_rows_kernel(const float* __restrict__ table, int64_t n_table, const int32_t* __restrict__ idx0,
                   const int32_t* __restrict__ level1, int n_inner, float* __restrict__ out, int64_t ldo,
                   int64_t rows, int Dq, int* __restrict__ err_flag) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * kGatherThreads + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kGatherThreads) >> 5;
    for (int64_t r = warp0; r < rows; r += nwarps) {
        int64_t src;
        if (level1 != nullptr) {
            const int64_t outer = r / n_inner, s = r % n_inner;
            src = level1[(int64_t)idx0[outer] * n_inner + s];
        } else {
            src = idx0[r];
        }
        if (src < 0 || src >= n_table) {          // the reference's index_select raises; we flag and read row 0
            if (lane == 0 && err_flag != nullptr) atomicExch(err_flag, 1);
            src = 0;
        }
        const float4* s4 = reinterpret_cast<const float4*>(table) + src * Dq;
        float4* d4 = reinterpret_cast<float4*>(out + r * ldo);
        for (int q = lane; q < Dq; q += 32) d4[q] = s4[q];
    }
}

inline int gather_grid(int64_t rows, int sm_count) {
    const int64_t warps_per_cta = kGatherThreads / 32;
    int64_t ctas = (rows + warps_per_cta - 1) / warps_per_cta;
    const int64_t cap = (int64_t)sm_count * 8;              // persistent-ish: 8 CTAs per SM, grid-stride over rows
    if (ctas > cap) ctas = cap;
    return (int)(ctas < 1 ? 1 : ctas);
}

// X_u[b] = [ history rows (gathered through hist_idx, or copied from `hist`) ; topic_emb ]
__global__ void __launch_bounds__(kGatherThreads)
build_user_nodes_kernel(const float* __restrict__ table, int64_t n_table, const int32_t* __restrict__ hist_idx,
                        const float* __restrict__ hist, const float* __restrict__ topic, float* __restrict__ Xu,
                        int64_t rows, int H, int C, int Dq, int* __restrict__ err_flag) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * kGatherThreads + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kGatherThreads) >> 5;
    const int nu = H + C;
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const int64_t b = r / nu;
        const int node = (int)(r % nu);
        const float4* s4;
        if (node >= H) {
            s4 = reinterpret_cast<const float4*>(topic) + (int64_t)(node - H) * Dq;
        } else if (table != nullptr) {
            int64_t src = hist_idx[b * H + node];
            if (src < 0 || src >= n_table) {
                if (lane == 0 && err_flag != nullptr) atomicExch(err_flag, 1);
                src = 0;
            }
            s4 = reinterpret_cast<const float4*>(table) + src * Dq;
        } else {
            s4 = reinterpret_cast<const float4*>(hist) + (b * H + node) * Dq;
        }
        float4* d4 = reinterpret_cast<float4*>(Xu) + r * Dq;
        for (int q = lane; q < Dq; q += 32) d4[q] = s4[q];
    }
}

__global__ void __launch_bounds__(kGatherThreads)
logits_kernel(const float* __restrict__ nc, const float* __restrict__ uc, float* __restrict__ out, int B, int Dq) {
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * kGatherThreads + threadIdx.x) >> 5;
    if (b >= B) return;
    const float4* n4 = reinterpret_cast<const float4*>(nc) + (size_t)b * Dq;
    const float4* u4 = reinterpret_cast<const float4*>(uc) + (size_t)b * Dq;
    float s = 0.f;
    for (int q = lane; q < Dq; q += 32) {
        const float4 x = n4[q], y = u4[q];
        s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
    }
    s = warp_sum(s);
    if (lane == 0) out[b] = s;
}

__global__ void add_inplace_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t count) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < count) {
        const float4 a = *reinterpret_cast<const float4*>(x + i);
        float4 b = *reinterpret_cast<float4*>(y + i);
        b.x += a.x; b.y += a.y; b.z += a.z; b.w += a.w;
        *reinterpret_cast<float4*>(y + i) = b;
    } else {
        for (int64_t k = i; k < count; ++k) y[k] += x[k];
    }
}

}  // namespace digat
