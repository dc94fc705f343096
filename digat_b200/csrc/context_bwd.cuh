// Backward kernels of the context ops (context.cuh): masked attention pooling and the topic-level segment
// softmax / segment sum.  Same CTA-per-batch-row layout as the forward kernels; HBM-bound.
#pragma once
#include "common.cuh"
#include "context.cuh"

namespace digat {

// block-wide sum of one float per thread (kCtxThreads threads); result broadcast to all threads
__device__ __forceinline__ float ctx_block_sum(float v, float* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = s_red[0];
#pragma unroll
    for (int w = 1; w < kCtxWarps; ++w) t += s_red[w];
    return t;
}

struct PoolBwdArgs {
    const float* F; int64_t strideF; int ldf;
    const float* resid;
    const float* v; const uint8_t* mask; const float* alpha;     // alpha [B,m] saved by the forward
    const float* dout; int ldg;                                   // [B, ldg]
    float* dF; float* dresid; float* dv;                          // dF/dresid dense [B,m,D]; dv [B,D]
    int B, m, D;
};

// out = sum_k alpha_k F'_k, alpha = softmax_k(mask(F'_k . v / sqrt(D))), F' = F or relu(F) + resid
//   dalpha_k = dout . F'_k ; t = sum_k alpha_k dalpha_k ; ds_k = mask_k ? alpha_k (dalpha_k - t) / sqrt(D) : 0
//   dF'_k = alpha_k dout + ds_k v ;  dv = sum_k ds_k F'_k ;  dF = dF' * 1[F > 0], dresid = dF' (when resid)
template <bool kResid>
__global__ void __launch_bounds__(kCtxThreads)
attention_pool_bwd_kernel(PoolBwdArgs p) {
    __shared__ float s_part[kCtxMaxItems][kCtxWarps];
    __shared__ float s_da[kCtxMaxItems];
    __shared__ float s_ds[kCtxMaxItems];
    __shared__ float s_al[kCtxMaxItems];
    __shared__ float s_red[kCtxWarps];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int m = p.m, D = p.D, nq = D >> 2;
    const float* F = p.F + (size_t)b * p.strideF;
    const float* Rs = kResid ? p.resid + (size_t)b * p.strideF : nullptr;
    const float* dout = p.dout + (size_t)b * p.ldg;
    ctx_scores<kResid>(F, p.ldf, Rs, p.ldf, dout, m, D, 1.0f, s_part, s_da);      // s_da[k] = dout . F'_k
    float mine = 0.f;
    if (tid < m) {
        s_al[tid] = p.alpha[(size_t)b * m + tid];
        mine = s_al[tid] * s_da[tid];
    }
    const float t = ctx_block_sum(mine, s_red);
    if (tid < m)
        s_ds[tid] = p.mask[(size_t)b * m + tid] != 0 ? s_al[tid] * (s_da[tid] - t) / sqrtf((float)D) : 0.f;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < kCtxMaxQuads; ++c) {
        const int q = tid + c * kCtxThreads;
        if (q >= nq) continue;
        const float4 g = reinterpret_cast<const float4*>(dout)[q];
        const float4 vv = reinterpret_cast<const float4*>(p.v + (size_t)b * D)[q];
        float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < m; ++k) {
            const float4 f = reinterpret_cast<const float4*>(F + (size_t)k * p.ldf)[q];
            float4 fp = f;
            if (kResid) {
                const float4 r = reinterpret_cast<const float4*>(Rs + (size_t)k * p.ldf)[q];
                fp.x = fmaxf(f.x, 0.f) + r.x; fp.y = fmaxf(f.y, 0.f) + r.y;
                fp.z = fmaxf(f.z, 0.f) + r.z; fp.w = fmaxf(f.w, 0.f) + r.w;
            }
            const float al = s_al[k], ds = s_ds[k];
            float4 d;
            d.x = fmaf(al, g.x, ds * vv.x); d.y = fmaf(al, g.y, ds * vv.y);
            d.z = fmaf(al, g.z, ds * vv.z); d.w = fmaf(al, g.w, ds * vv.w);
            dv.x = fmaf(ds, fp.x, dv.x); dv.y = fmaf(ds, fp.y, dv.y);
            dv.z = fmaf(ds, fp.z, dv.z); dv.w = fmaf(ds, fp.w, dv.w);
            const size_t o = ((size_t)b * m + k) * D;
            if (kResid) {
                reinterpret_cast<float4*>(p.dresid + o)[q] = d;
                d.x = f.x > 0.f ? d.x : 0.f; d.y = f.y > 0.f ? d.y : 0.f;
                d.z = f.z > 0.f ? d.z : 0.f; d.w = f.w > 0.f ? d.w : 0.f;
            }
            reinterpret_cast<float4*>(p.dF + o)[q] = d;
        }
        reinterpret_cast<float4*>(p.dv + (size_t)b * D)[q] = dv;
    }
}

inline int launch_attention_pool_bwd(const float* F, int64_t strideF, int ldf, const float* resid, const float* v,
                                     const uint8_t* mask, const float* alpha, const float* dout, int ldg, float* dF,
                                     float* dresid, float* dv, int B, int m, int D, cudaStream_t st) {
    DIGAT_REQUIRE(F && v && mask && alpha && dout && dF && dv && (!resid || dresid), "digat_attention_pool_bwd: null pointer");
    DIGAT_REQUIRE(m >= 1 && m <= kCtxMaxItems, "digat_attention_pool_bwd: m=%d outside [1,%d]", m, kCtxMaxItems);
    DIGAT_REQUIRE(D >= 4 && (D & 3) == 0 && D <= 4 * kCtxThreads * kCtxMaxQuads, "digat_attention_pool_bwd: bad D=%d", D);
    DIGAT_REQUIRE((ldf & 3) == 0 && (ldg & 3) == 0 && (strideF & 3) == 0 && ldf >= D && ldg >= D,
                  "digat_attention_pool_bwd: strides must be multiples of 4 and >= D");
    DIGAT_REQUIRE(aligned16(F) && aligned16(v) && aligned16(dout) && aligned16(dF) && aligned16(dv) &&
                  (!resid || (aligned16(resid) && aligned16(dresid))), "digat_attention_pool_bwd: pointers must be 16-byte aligned");
    if (B <= 0) return DIGAT_OK;
    PoolBwdArgs a{F, strideF, ldf, resid, v, mask, alpha, dout, ldg, dF, dresid, dv, B, m, D};
    if (resid) attention_pool_bwd_kernel<true><<<B, kCtxThreads, 0, st>>>(a);
    else       attention_pool_bwd_kernel<false><<<B, kCtxThreads, 0, st>>>(a);
    return check_launch("digat_attention_pool_bwd");
}

struct SegBwdArgs {
    const float* Xu; int64_t strideX;
    const float* v; const int64_t* cidx; const float* alpha;     // alpha [B,H] saved by the forward
    const float* dT;                                              // [B, n_seg, D]
    float* dXu;                                                   // [B, n_u, D] (rows >= H are zero-filled)
    float* dv;                                                    // [B, D]
    int B, H, n_seg, n_u, D;
};

// T_k = sum_{t in seg k} alpha_t Xh_t, alpha = segment softmax of a_t = Xh_t . v / sqrt(D)
//   dalpha_t = dT_{c(t)} . Xh_t ; tt_k = sum_{t in k} alpha_t dalpha_t ; ds_t = alpha_t (dalpha_t - tt_{c(t)}) / sqrt(D)
//   dXh_t = alpha_t dT_{c(t)} + ds_t v ;  dv = sum_t ds_t Xh_t
__global__ void __launch_bounds__(kCtxThreads)
topic_segment_bwd_kernel(SegBwdArgs p) {
    // dalpha_t and the segment means are kept in DOUBLE: ds_t = alpha_t (dalpha_t - sum_k alpha_k dalpha_k) cancels when the
    // history rows of a segment are alike (after smoothing graph layers they are), and an fp32 dot product's 1e-7 then shows up
    // as 3e-5 in the gradients of user_news_K / user_news_Q (measured on the vanilla-GAT ablation, torch's fp32 autograd: 1e-6)
    __shared__ double s_part[kCtxMaxItems][kCtxWarps];
    __shared__ double s_da[kCtxMaxItems];
    __shared__ float s_ds[kCtxMaxItems];
    __shared__ float s_al[kCtxMaxItems];
    __shared__ int s_seg[kCtxMaxItems];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, D = p.D, nq = D >> 2, n_seg = p.n_seg;
    const float* Xh = p.Xu + (size_t)b * p.strideX;
    const float* dT = p.dT + (size_t)b * n_seg * D;
    if (tid < H) {
        const int64_t c = p.cidx[(size_t)b * H + tid];
        s_seg[tid] = (c < 0 || c >= n_seg) ? n_seg - 1 : (int)c;
        s_al[tid] = p.alpha[(size_t)b * H + tid];
    }
    __syncthreads();
    // Rows are handled five at a time: their loads are independent, so one memory round trip covers five rows (a training
    // batch is a few hundred CTAs -- the kernel is latency-bound, and one row per trip cost 47 us for 25 MB).
    constexpr int kRows = 5;
    for (int t0 = 0; t0 < H; t0 += kRows) {
        float4 x[kRows][kCtxMaxQuads], d[kRows][kCtxMaxQuads];
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
            const int t = min(t0 + u, H - 1);                                  // the tail repeats the last row (result unused)
            const float* drow = dT + (size_t)s_seg[t] * D;
#pragma unroll
            for (int c = 0; c < kCtxMaxQuads; ++c) {
                const int q = tid + c * kCtxThreads;
                x[u][c] = q < nq ? reinterpret_cast<const float4*>(Xh + (size_t)t * D)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                d[u][c] = q < nq ? reinterpret_cast<const float4*>(drow)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        double part[kRows];
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
            part[u] = 0.0;
#pragma unroll
            for (int c = 0; c < kCtxMaxQuads; ++c) {
                part[u] += (double)x[u][c].x * d[u][c].x + (double)x[u][c].y * d[u][c].y +
                           (double)x[u][c].z * d[u][c].z + (double)x[u][c].w * d[u][c].w;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < kRows; ++u) part[u] += __shfl_xor_sync(0xffffffffu, part[u], o);
        if (lane == 0) {
#pragma unroll
            for (int u = 0; u < kRows; ++u)
                if (t0 + u < H) s_part[t0 + u][warp] = part[u];
        }
    }
    __syncthreads();
    if (tid < H) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kCtxWarps; ++w) s += s_part[tid][w];
        s_da[tid] = s;
    }
    __syncthreads();
    if (tid < H) {
        const int me = s_seg[tid];
        double tt = 0.0;
        for (int t = 0; t < H; ++t)
            if (s_seg[t] == me) tt += (double)s_al[t] * s_da[t];
        s_ds[tid] = (float)((double)s_al[tid] * (s_da[tid] - tt) / sqrt((double)D));
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < kCtxMaxQuads; ++c) {
        const int q = tid + c * kCtxThreads;
        if (q >= nq) continue;
        const float4 vv = reinterpret_cast<const float4*>(p.v + (size_t)b * D)[q];
        float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
        float* dX = p.dXu + (size_t)b * p.n_u * D;
        for (int t0 = 0; t0 < H; t0 += kRows) {
            float4 x[kRows], d[kRows];
#pragma unroll
            for (int u = 0; u < kRows; ++u) {
                const int t = min(t0 + u, H - 1);
                x[u] = reinterpret_cast<const float4*>(Xh + (size_t)t * D)[q];
                d[u] = reinterpret_cast<const float4*>(dT + (size_t)s_seg[t] * D)[q];
            }
#pragma unroll
            for (int u = 0; u < kRows; ++u) {
                const int t = t0 + u;
                if (t >= H) break;
                const float al = s_al[t], ds = s_ds[t];
                float4 o;
                o.x = fmaf(al, d[u].x, ds * vv.x); o.y = fmaf(al, d[u].y, ds * vv.y);
                o.z = fmaf(al, d[u].z, ds * vv.z); o.w = fmaf(al, d[u].w, ds * vv.w);
                dv.x = fmaf(ds, x[u].x, dv.x); dv.y = fmaf(ds, x[u].y, dv.y);
                dv.z = fmaf(ds, x[u].z, dv.z); dv.w = fmaf(ds, x[u].w, dv.w);
                reinterpret_cast<float4*>(dX + (size_t)t * D)[q] = o;
            }
        }
        for (int t = H; t < p.n_u; ++t)
            reinterpret_cast<float4*>(dX + (size_t)t * D)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(p.dv + (size_t)b * D)[q] = dv;
    }
}

inline int launch_topic_segment_bwd(const float* Xu, int64_t strideX, const float* v, const int64_t* cidx,
                                    const float* alpha, const float* dT, float* dXu, float* dv, int B, int H, int n_seg,
                                    int n_u, int D, cudaStream_t st) {
    DIGAT_REQUIRE(Xu && v && cidx && alpha && dT && dXu && dv, "digat_topic_segment_bwd: null pointer");
    DIGAT_REQUIRE(H >= 1 && H <= kCtxMaxItems && n_seg >= 1 && n_seg <= kCtxMaxItems && n_u >= H,
                  "digat_topic_segment_bwd: bad H / n_seg / n_u");
    DIGAT_REQUIRE(D >= 4 && (D & 3) == 0 && D <= 4 * kCtxThreads * kCtxMaxQuads, "digat_topic_segment_bwd: bad D=%d", D);
    DIGAT_REQUIRE((strideX & 3) == 0 && aligned16(Xu) && aligned16(v) && aligned16(dT) && aligned16(dXu) && aligned16(dv),
                  "digat_topic_segment_bwd: pointers/strides must be 16-byte aligned");
    if (B <= 0) return DIGAT_OK;
    SegBwdArgs a{Xu, strideX, v, cidx, alpha, dT, dXu, dv, B, H, n_seg, n_u, D};
    topic_segment_bwd_kernel<<<B, kCtxThreads, 0, st>>>(a);
    return check_launch("digat_topic_segment_bwd");
}

// ---------------------------------------------------------------------------------------------- news gate backward
// out = ctx_in + s l + (1 - s) g with s = sigmoid(z)  (graphEncoders.py:112-113):
//   dz = dout (l - g) s (1 - s),  dl = dout s,  dg = dout (1 - s);  d ctx_in = dout (the caller passes it on).
__global__ void news_gate_bwd_kernel(const float4* __restrict__ z, const float4* __restrict__ lg, const float4* __restrict__ dout,
                                     float4* __restrict__ dz, float4* __restrict__ dlg, int B, int Dq) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * Dq) return;
    const int b = (int)(i / Dq), q = (int)(i % Dq);
    const float4 zz = z[i], d = dout[i];
    const float4 l = lg[(size_t)b * 2 * Dq + q];
    const float4 g = lg[(size_t)b * 2 * Dq + Dq + q];
    float4 oz, ol, og;
    auto one = [](float zv, float lv, float gv, float dv, float& z_, float& l_, float& g_) {
        const float s = 1.f / (1.f + expf(-zv));
        z_ = dv * (lv - gv) * (s * (1.f - s));
        l_ = dv * s;
        g_ = dv * (1.f - s);
    };
    one(zz.x, l.x, g.x, d.x, oz.x, ol.x, og.x); one(zz.y, l.y, g.y, d.y, oz.y, ol.y, og.y);
    one(zz.z, l.z, g.z, d.z, oz.z, ol.z, og.z); one(zz.w, l.w, g.w, d.w, oz.w, ol.w, og.w);
    dz[i] = oz;
    dlg[(size_t)b * 2 * Dq + q] = ol;
    dlg[(size_t)b * 2 * Dq + Dq + q] = og;
}

inline int launch_news_gate_bwd(const float* z, const float* lg, const float* dout, float* dz, float* dlg, int B, int D,
                                cudaStream_t st) {
    if (B <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(z && lg && dout && dz && dlg, "digat_news_gate_bwd: null pointer");
    DIGAT_REQUIRE(D >= 4 && (D & 3) == 0, "digat_news_gate_bwd: D must be a multiple of 4");
    DIGAT_REQUIRE(aligned16(z) && aligned16(lg) && aligned16(dout) && aligned16(dz) && aligned16(dlg),
                  "digat_news_gate_bwd: pointers must be 16-byte aligned");
    const int64_t total = (int64_t)B * (D / 4);
    news_gate_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(z), reinterpret_cast<const float4*>(lg), reinterpret_cast<const float4*>(dout),
        reinterpret_cast<float4*>(dz), reinterpret_cast<float4*>(dlg), B, D / 4);
    return check_launch("digat_news_gate_bwd");
}

}  // namespace digat
