// Clip-by-global-norm + Adam over FLAT fp32 buffers: the optimizer step of reference trainer.py:98-105
// (nn.utils.clip_grad_norm_ then optim.Adam.step) as two launches instead of PyTorch's ~150 per-parameter kernels
// (with capturable=True the bias-correction divisions fall off the foreach fast path: 2 launches per parameter).
//   1. sumsq_partials_kernel: fixed 1024 partial sums of g^2 (block b owns a contiguous chunk: deterministic);
//      thread 0 of block 0 advances the device-side step counter (CUDA-graph replays need it on the device)
//   2. adam_clip_kernel: every block re-reduces the 1024 partials in the same order (double), derives the clip
//      coefficient min(1, max_norm / (norm + 1e-6)) and the bias corrections, then updates its elements:
//        g <- g * coef (written back, as clip_grad_norm_ does);  g' = g + wd * p
//        m <- m + (g' - m)(1 - b1);  v <- v b2 + (1 - b2) g'^2;  p <- p - lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#pragma once
#include "common.cuh"

namespace digat {

constexpr int kOptPartials = 1024, kOptThreads = 256;

__global__ void __launch_bounds__(kOptThreads)
sumsq_partials_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partials, float* __restrict__ step) {
    __shared__ float red[kOptThreads / 32];
    const int64_t nq = n >> 2;                                              // float4 body; the tail (< 4) goes to the last block
    const int64_t per = (nq + kOptPartials - 1) / kOptPartials;
    const int64_t lo = (int64_t)blockIdx.x * per, hi = lo + per < nq ? lo + per : nq;
    float s = 0.f;
    for (int64_t i = lo + threadIdx.x; i < hi; i += kOptThreads) {
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    }
    if (blockIdx.x == kOptPartials - 1 && threadIdx.x == 0)
        for (int64_t i = nq << 2; i < n; ++i) s = fmaf(g[i], g[i], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kOptThreads / 32; ++w) t += red[w];
        partials[blockIdx.x] = t;
        if (blockIdx.x == 0 && step != nullptr) step[0] += 1.f;
    }
}

struct AdamArgs {
    float* p; float* g; float* m; float* v; int64_t n;
    const float* partials; const float* step; float* norm_out;
    float max_norm, lr, beta1, beta2, eps, weight_decay;
};

__global__ void __launch_bounds__(kOptThreads)
adam_clip_kernel(AdamArgs a) {
    __shared__ double red[kOptThreads / 32];
    __shared__ float s_coef;
    double t = 0.0;
    for (int i = threadIdx.x; i < kOptPartials; i += kOptThreads) t += (double)a.partials[i];   // 4 per thread, fixed order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < kOptThreads / 32; ++w) tot += red[w];
        const float norm = (float)sqrt(tot);
        float coef = 1.f;
        if (a.max_norm > 0.f) {
            coef = a.max_norm / (norm + 1e-6f);
            coef = coef < 1.f ? coef : 1.f;
        }
        s_coef = coef;
        if (blockIdx.x == 0 && a.norm_out != nullptr) a.norm_out[0] = norm;
    }
    __syncthreads();
    const float coef = s_coef;
    const double steps = (double)a.step[0];
    const float bc1 = (float)(1.0 - pow((double)a.beta1, steps));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)a.beta2, steps));
    const float step_size = a.lr / bc1;
    const float w1 = 1.f - a.beta1, w2 = 1.f - a.beta2;
    auto upd = [&](float& p, float& g, float& m, float& v) {
        g *= coef;
        const float gg = a.weight_decay != 0.f ? fmaf(a.weight_decay, p, g) : g;
        m = fmaf(gg - m, w1, m);
        v = fmaf(w2 * gg, gg, v * a.beta2);
        p -= step_size * (m / (sqrtf(v) / bc2_sqrt + a.eps));
    };
    const int64_t nq = a.n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < nq; i += (int64_t)gridDim.x * kOptThreads) {
        float4 p = reinterpret_cast<float4*>(a.p)[i], g = reinterpret_cast<float4*>(a.g)[i];
        float4 m = reinterpret_cast<float4*>(a.m)[i], v = reinterpret_cast<float4*>(a.v)[i];
        upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
        reinterpret_cast<float4*>(a.p)[i] = p; reinterpret_cast<float4*>(a.g)[i] = g;
        reinterpret_cast<float4*>(a.m)[i] = m; reinterpret_cast<float4*>(a.v)[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(a.n & 3)) {
        const int64_t i = (nq << 2) + threadIdx.x;
        upd(a.p[i], a.g[i], a.m[i], a.v[i]);
    }
}

inline int launch_grad_sumsq(const float* g, int64_t n, float* partials, float* step, cudaStream_t st) {
    DIGAT_REQUIRE(g && partials && n > 0 && aligned16(g), "digat_grad_sumsq: null / misaligned pointer or n <= 0");
    sumsq_partials_kernel<<<kOptPartials, kOptThreads, 0, st>>>(g, n, partials, step);
    return check_launch("digat_grad_sumsq");
}

inline int launch_adam_clip_step(float* p, float* g, float* m, float* v, int64_t n, const float* partials, const float* step,
                                 float* norm_out, float max_norm, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, cudaStream_t st) {
    DIGAT_REQUIRE(p && g && m && v && partials && step && n > 0, "digat_adam_clip_step: null pointer or n <= 0");
    DIGAT_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), "digat_adam_clip_step: buffers must be 16-byte aligned");
    DIGAT_REQUIRE(lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f,
                  "digat_adam_clip_step: bad hyper-parameter");
    const DeviceInfo* di = device_info();
    if (!di) return fail(DIGAT_E_CUDA, "digat_adam_clip_step: no CUDA device");
    const int64_t blocks = ((n >> 2) + kOptThreads - 1) / kOptThreads;
    const int64_t cap = (int64_t)8 * di->sm_count;
    AdamArgs a{p, g, m, v, n, partials, step, norm_out, max_norm, lr, beta1, beta2, eps, weight_decay};
    adam_clip_kernel<<<(unsigned)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap)), kOptThreads, 0, st>>>(a);
    return check_launch("digat_adam_clip_step");
}

}  // namespace digat
