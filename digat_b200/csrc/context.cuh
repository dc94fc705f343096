// Context kernels of the DIGAT encoder: masked single-query attention pooling (reference layers.py:199-206),
// the news-graph gate (graphEncoders.py:112-113) and the topic-level segment softmax / segment sum of the user
// history (graphEncoders.py:128-130, torch_scatter.scatter_softmax + scatter_sum).
//
// All three are HBM-bound: every feature row is read once from DRAM (the value pass re-reads it from L1/L2), so
// the roofline is bytes(F) / HBM bandwidth.  One CTA of kCtxThreads threads per batch row; a thread owns one or
// more float4 feature quads, dot products are reduced with warp shuffles + one smem hop.
#pragma once
#include "common.cuh"

namespace digat {

constexpr int kCtxThreads = 128;
constexpr int kCtxWarps = kCtxThreads / 32;
constexpr int kCtxMaxItems = 128;     // max features per pooled set (graph nodes / topics / history length)
constexpr int kCtxMaxQuads = 2;       // D <= 4 * kCtxThreads * kCtxMaxQuads = 1024

// scores[k] = (F'_k . v) / sqrt(D) for k in [0,m), F' = F or relu(F)+resid.  Result in s_score (smem, all threads sync'd).
// Rows are processed four at a time so that four rows' loads are in flight per thread (the loop is latency-bound
// otherwise: one 1600-byte row per DRAM round trip).  s_skip (smem, may be null): rows with s_skip[k] != 0 are not
// read at all (their score is 0; the caller masks them).
template <bool kResid>
__device__ __forceinline__ void ctx_scores(const float* __restrict__ F, int ldf, const float* __restrict__ Rs, int ldr,
                                           const float* __restrict__ v, int m, int D, float inv_scale_div,
                                           float (*s_part)[kCtxWarps], float* s_score, const uint8_t* s_skip = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nq = D >> 2;
    float4 vq[kCtxMaxQuads];
#pragma unroll
    for (int c = 0; c < kCtxMaxQuads; ++c) {
        const int q = tid + c * kCtxThreads;
        vq[c] = q < nq ? reinterpret_cast<const float4*>(v)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    constexpr int kRows = 4;
    for (int k0 = 0; k0 < m; k0 += kRows) {
        float4 f[kRows][kCtxMaxQuads], t[kRows][kCtxMaxQuads];
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
            const int k = k0 + u;
            const bool live = k < m && !(s_skip != nullptr && s_skip[k] != 0);
#pragma unroll
            for (int c = 0; c < kCtxMaxQuads; ++c) {
                const int q = tid + c * kCtxThreads;
                const bool on = live && q < nq;
                f[u][c] = on ? reinterpret_cast<const float4*>(F + (size_t)k * ldf)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                if (kResid)
                    t[u][c] = on ? reinterpret_cast<const float4*>(Rs + (size_t)k * ldr)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        float part[kRows];
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
            part[u] = 0.f;
#pragma unroll
            for (int c = 0; c < kCtxMaxQuads; ++c) {
                float4 x = f[u][c];
                if (kResid) {
                    x.x = fmaxf(x.x, 0.f) + t[u][c].x; x.y = fmaxf(x.y, 0.f) + t[u][c].y;
                    x.z = fmaxf(x.z, 0.f) + t[u][c].z; x.w = fmaxf(x.w, 0.f) + t[u][c].w;
                }
                part[u] = fmaf(x.x, vq[c].x, part[u]);
                part[u] = fmaf(x.y, vq[c].y, part[u]);
                part[u] = fmaf(x.z, vq[c].z, part[u]);
                part[u] = fmaf(x.w, vq[c].w, part[u]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < kRows; ++u) part[u] += __shfl_xor_sync(0xffffffffu, part[u], o);
        if (lane == 0) {
#pragma unroll
            for (int u = 0; u < kRows; ++u)
                if (k0 + u < m) s_part[k0 + u][warp] = part[u];
        }
    }
    __syncthreads();
    if (tid < m) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kCtxWarps; ++w) s += s_part[tid][w];
        s_score[tid] = s / inv_scale_div;
    }
    __syncthreads();
}

struct PoolArgs {
    const float* F; int64_t strideF; int ldf;
    const float* resid;                 // same strides as F (may be null)
    const float* v; int ldv;            // [B, ldv] query-side vector (row b at v + b*ldv)
    const uint8_t* mask;                // [B, m]
    float* out; int ldo;                // [B, ldo]
    const float* add_in;                // optional [B, ldo]: out = add_in + pooled (may alias out)
    float* first_out;                   // optional [B, ldo]: copy of F[b, 0, :]
    float* alpha_out;                   // optional [B, m]
    int B, m, D;
};

template <bool kResid>
__global__ void __launch_bounds__(kCtxThreads)
attention_pool_fwd_kernel(PoolArgs p) {
    __shared__ float s_part[kCtxMaxItems][kCtxWarps];
    __shared__ float s_score[kCtxMaxItems];
    __shared__ float s_red[kCtxWarps];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = p.m, D = p.D, nq = D >> 2;
    const float* F = p.F + (size_t)b * p.strideF;
    const float* Rs = kResid ? p.resid + (size_t)b * p.strideF : nullptr;
    // Masked rows get the -1e9 fill whatever their score is, and a softmax weight of exactly 0 unless EVERY row is
    // masked (then the weights are uniform): they are not read at all when at least one row is unmasked.
    __shared__ uint8_t s_skip[kCtxMaxItems];
    const bool unmasked = tid < m && p.mask[(size_t)b * m + tid] != 0;
    const int any_unmasked = __syncthreads_or(unmasked ? 1 : 0);
    if (tid < m) s_skip[tid] = (any_unmasked && !unmasked) ? 1 : 0;
    __syncthreads();
    ctx_scores<kResid>(F, p.ldf, Rs, p.ldf, p.v + (size_t)b * p.ldv, m, D, sqrtf((float)D), s_part, s_score, s_skip);

    // masked softmax over the m scores (m <= 128 = one value per thread)
    float val = -INFINITY;
    if (tid < m) val = p.mask[(size_t)b * m + tid] != 0 ? s_score[tid] : kNegFill;
    float mx = warp_max(val);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = s_red[0];
#pragma unroll
    for (int w = 1; w < kCtxWarps; ++w) mx = fmaxf(mx, s_red[w]);
    __syncthreads();
    const float e = tid < m ? expf(val - mx) : 0.f;
    float sum = warp_sum(e);
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    sum = s_red[0];
#pragma unroll
    for (int w = 1; w < kCtxWarps; ++w) sum += s_red[w];
    if (tid < m) {
        const float al = e / sum;
        s_score[tid] = al;
        if (p.alpha_out != nullptr) p.alpha_out[(size_t)b * m + tid] = al;
    }
    __syncthreads();

    // value pass: out = sum_k alpha_k F'_k  (k ascending, like bmm's reduction over the feature axis)
#pragma unroll
    for (int c = 0; c < kCtxMaxQuads; ++c) {
        const int q = tid + c * kCtxThreads;
        if (q >= nq) continue;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        constexpr int kRows = 4;                                   // four rows' loads in flight per thread
        for (int k0 = 0; k0 < m; k0 += kRows) {
            float4 f[kRows], t[kRows];
            float al[kRows];
#pragma unroll
            for (int u = 0; u < kRows; ++u) {
                const int k = k0 + u;
                al[u] = k < m ? s_score[k] : 0.f;
                const bool live = k < m && s_skip[k] == 0;         // skipped rows have weight 0: fma(0, 0, acc) == acc
                f[u] = live ? reinterpret_cast<const float4*>(F + (size_t)k * p.ldf)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                if (kResid)
                    t[u] = live ? reinterpret_cast<const float4*>(Rs + (size_t)k * p.ldf)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kRows; ++u) {
                float4 x = f[u];
                if (kResid) {
                    x.x = fmaxf(x.x, 0.f) + t[u].x; x.y = fmaxf(x.y, 0.f) + t[u].y;
                    x.z = fmaxf(x.z, 0.f) + t[u].z; x.w = fmaxf(x.w, 0.f) + t[u].w;
                }
                acc.x = fmaf(al[u], x.x, acc.x); acc.y = fmaf(al[u], x.y, acc.y);
                acc.z = fmaf(al[u], x.z, acc.z); acc.w = fmaf(al[u], x.w, acc.w);
            }
        }
        if (p.add_in != nullptr) {
            const float4 c = reinterpret_cast<const float4*>(p.add_in + (size_t)b * p.ldo)[q];
            acc.x = c.x + acc.x; acc.y = c.y + acc.y; acc.z = c.z + acc.z; acc.w = c.w + acc.w;
        }
        reinterpret_cast<float4*>(p.out + (size_t)b * p.ldo)[q] = acc;
        if (p.first_out != nullptr)
            reinterpret_cast<float4*>(p.first_out + (size_t)b * p.ldo)[q] = reinterpret_cast<const float4*>(F)[q];
    }
}

inline int launch_attention_pool_fwd(const float* F, int64_t strideF, int ldf, const float* resid, const float* v,
                                     int ldv, const uint8_t* mask, const float* add_in, float* out, int ldo, float* first_out,
                                     float* alpha_out, int B, int m, int D, cudaStream_t st) {
    if (B <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(F && v && mask && out, "digat_attention_pool_fwd: null pointer");
    DIGAT_REQUIRE(m >= 1 && m <= kCtxMaxItems, "digat_attention_pool_fwd: m=%d outside [1,%d]", m, kCtxMaxItems);
    DIGAT_REQUIRE(D >= 4 && (D & 3) == 0 && D <= 4 * kCtxThreads * kCtxMaxQuads, "digat_attention_pool_fwd: bad D=%d", D);
    DIGAT_REQUIRE((ldf & 3) == 0 && (ldo & 3) == 0 && (strideF & 3) == 0 && ldf >= D && ldo >= D,
                  "digat_attention_pool_fwd: strides must be multiples of 4 and >= D");
    DIGAT_REQUIRE(aligned16(F) && aligned16(v) && aligned16(out) && (!resid || aligned16(resid)) &&
                  (!first_out || aligned16(first_out)), "digat_attention_pool_fwd: pointers must be 16-byte aligned");
    if (B <= 0) return DIGAT_OK;
    DIGAT_REQUIRE((ldv & 3) == 0 && ldv >= D, "digat_attention_pool_fwd: ldv must be a multiple of 4 and >= D");
    PoolArgs a{F, strideF, ldf, resid, v, ldv, mask, out, ldo, add_in, first_out, alpha_out, B, m, D};
    if (resid) attention_pool_fwd_kernel<true><<<B, kCtxThreads, 0, st>>>(a);
    else       attention_pool_fwd_kernel<false><<<B, kCtxThreads, 0, st>>>(a);
    return check_launch("digat_attention_pool_fwd");
}

// ---------------------------------------------------------------------------------------------- news gate
__global__ void news_gate_fwd_kernel(const float4* __restrict__ z, const float4* __restrict__ lg,
                                     const float4* ctx_in, float4* ctx_out, int B, int Dq) {   // ctx_in may alias ctx_out
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * Dq) return;
    const int b = (int)(i / Dq), q = (int)(i % Dq);
    const float4 zz = z[i];
    const float4 l = lg[(size_t)b * 2 * Dq + q];
    const float4 g = lg[(size_t)b * 2 * Dq + Dq + q];
    float4 o = ctx_in != nullptr ? ctx_in[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    auto mix = [](float zv, float lv, float gv) {
        const float gate = 1.f / (1.f + expf(-zv));
        return gate * lv + (1.f - gate) * gv;
    };
    o.x += mix(zz.x, l.x, g.x); o.y += mix(zz.y, l.y, g.y);
    o.z += mix(zz.z, l.z, g.z); o.w += mix(zz.w, l.w, g.w);
    ctx_out[i] = o;
}

inline int launch_news_gate_fwd(const float* z, const float* lg, const float* ctx_in, float* ctx_out, int B, int D,
                                cudaStream_t st) {
    if (B <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(z && lg && ctx_out, "digat_news_gate_fwd: null pointer");
    DIGAT_REQUIRE(D >= 4 && (D & 3) == 0, "digat_news_gate_fwd: D must be a multiple of 4");
    DIGAT_REQUIRE(aligned16(z) && aligned16(lg) && aligned16(ctx_out) && (!ctx_in || aligned16(ctx_in)),
                  "digat_news_gate_fwd: pointers must be 16-byte aligned");
    if (B <= 0) return DIGAT_OK;
    const int64_t total = (int64_t)B * (D / 4);
    news_gate_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(z), reinterpret_cast<const float4*>(lg),
        reinterpret_cast<const float4*>(ctx_in), reinterpret_cast<float4*>(ctx_out), B, D / 4);
    return check_launch("digat_news_gate_fwd");
}

// ---------------------------------------------------------------------------------------------- topic segments
struct SegArgs {
    const float* Xu; int64_t strideX;     // [B, n_u, D], first H rows are the history
    const float* v; int ldv;              // [B, ldv]
    const int64_t* cidx;                  // [B, H]
    float* T;                             // [B, n_seg, D]
    float* alpha_out;                     // optional [B, H]
    int B, H, n_seg, D;
    int* err_flag;                        // device int, set to 1 when a segment id is out of range
    const int32_t* src_index;             // optional [B]: row b reads Xu and cidx of row src_index[b] (shared user graphs)
    const uint8_t* cmask;                 // optional [B, n_seg]: segments masked out of the user-level attention are not
                                          // evaluated (T = 0) unless every segment of the row is masked
    float* Tc; const int32_t* seg_pos;    // optional (with cmask): the evaluated segments are ALSO written to the compact
                                          // list Tc[seg_pos[b*n_seg + k]] -- the operand of the pruned featureAffine GEMM
};

// One CTA per (user, candidate) row.  The kernel is organised around a compact, segment-sorted list of the LIVE history
// slots (slots of segments the user-level attention masks out are dropped up front): both streaming passes walk that
// list four rows at a time, so per-row control work stays small next to the 1600-byte row it loads.
__global__ void __launch_bounds__(kCtxThreads)
topic_segment_fwd_kernel(SegArgs p) {
    __shared__ float s_part[kCtxMaxItems][kCtxWarps];
    __shared__ float s_score[kCtxMaxItems];    // per slot: score, then exp(score - segment max)
    __shared__ float s_alpha[kCtxMaxItems];
    __shared__ int s_seg[kCtxMaxItems];
    __shared__ int s_order[kCtxMaxItems];      // live history slots sorted by (segment, slot): a stable counting sort
    __shared__ int s_start[kCtxMaxItems + 1];  // segment k owns s_order[s_start[k] .. s_start[k+1])
    __shared__ uint8_t s_segskip[kCtxMaxItems];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, D = p.D, nq = D >> 2, n_seg = p.n_seg;
    const size_t src = p.src_index != nullptr ? (size_t)p.src_index[b] : (size_t)b;
    const float* Xh = p.Xu + src * p.strideX;
    if (tid < H) {
        const int64_t c = p.cidx[src * H + tid];
        int ci = (int)c;
        if (c < 0 || c >= n_seg) { ci = n_seg - 1; if (p.err_flag) atomicExch(p.err_flag, 1); }
        s_seg[tid] = ci;
    }
    {
        const bool unmasked = p.cmask != nullptr && tid < n_seg && p.cmask[(size_t)b * n_seg + tid] != 0;
        const int any_unmasked = __syncthreads_or(unmasked ? 1 : 0);           // (also orders the s_seg writes)
        if (tid < n_seg) s_segskip[tid] = (p.cmask != nullptr && any_unmasked && !unmasked) ? 1 : 0;
        __syncthreads();
    }
    // stable counting sort of the live slots by segment (thread k handles segment k)
    for (int k = tid; k <= n_seg; k += kCtxThreads) {
        int cnt = 0;
        for (int t = 0; t < H; ++t) {
            const int sg = s_seg[t];
            cnt += (sg < k && s_segskip[sg] == 0);
        }
        s_start[k] = cnt;
    }
    __syncthreads();
    for (int k = tid; k < n_seg; k += kCtxThreads) {
        if (s_segskip[k]) continue;
        int pos = s_start[k];
        for (int t = 0; t < H; ++t)
            if (s_seg[t] == k) s_order[pos++] = t;
    }
    __syncthreads();
    const int n_live = s_start[n_seg];

    // ---- pass 1: scores of the live slots, a_t = Xh_t . v / sqrt(D)
    float4 vq[kCtxMaxQuads];
#pragma unroll
    for (int c = 0; c < kCtxMaxQuads; ++c) {
        const int q = tid + c * kCtxThreads;
        vq[c] = q < nq ? reinterpret_cast<const float4*>(p.v + (size_t)b * p.ldv)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    constexpr int kRows = 4;
    for (int e0 = 0; e0 < n_live; e0 += kRows) {
        float4 x[kRows][kCtxMaxQuads];
        int tt[kRows];
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
            tt[u] = s_order[min(e0 + u, n_live - 1)];                          // the tail repeats the last row (result unused)
#pragma unroll
            for (int c = 0; c < kCtxMaxQuads; ++c) {
                const int q = tid + c * kCtxThreads;
                x[u][c] = q < nq ? reinterpret_cast<const float4*>(Xh + (size_t)tt[u] * D)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        float part[kRows];
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
            part[u] = 0.f;
#pragma unroll
            for (int c = 0; c < kCtxMaxQuads; ++c) {
                part[u] = fmaf(x[u][c].x, vq[c].x, part[u]);
                part[u] = fmaf(x[u][c].y, vq[c].y, part[u]);
                part[u] = fmaf(x[u][c].z, vq[c].z, part[u]);
                part[u] = fmaf(x[u][c].w, vq[c].w, part[u]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < kRows; ++u) part[u] += __shfl_xor_sync(0xffffffffu, part[u], o);
        if (lane == 0) {
#pragma unroll
            for (int u = 0; u < kRows; ++u)
                if (e0 + u < n_live) s_part[tt[u]][warp] = part[u];
        }
    }
    __syncthreads();
    const bool live_slot = tid < H && s_segskip[s_seg[tid]] == 0;
    const float inv_div = sqrtf((float)D);
    if (live_slot) {
        float sc = 0.f;
#pragma unroll
        for (int w = 0; w < kCtxWarps; ++w) sc += s_part[tid][w];
        s_score[tid] = sc / inv_div;
    }
    __syncthreads();
    // ---- segment softmax over the slot's own segment (its sorted range, ascending slots): max, exp once per slot, sum
    int r0 = 0, r1 = 0;
    float mx = 0.f;
    if (live_slot) {
        r0 = s_start[s_seg[tid]];
        r1 = s_start[s_seg[tid] + 1];
        mx = -INFINITY;
        for (int e = r0; e < r1; ++e) mx = fmaxf(mx, s_score[s_order[e]]);
    }
    __syncthreads();
    if (live_slot) s_score[tid] = expf(s_score[tid] - mx);
    __syncthreads();
    if (tid < H) {
        float al = 0.f;
        if (live_slot) {
            float sum = 0.f;
            for (int e = r0; e < r1; ++e) sum += s_score[s_order[e]];
            al = s_score[tid] / sum;
        }
        s_alpha[tid] = al;
        if (p.alpha_out != nullptr) p.alpha_out[(size_t)b * H + tid] = al;
    }
    __syncthreads();

    // ---- pass 2: T[k] = sum over segment k (ascending slots) of alpha_t * Xh_t; segments without live slots are 0
#pragma unroll
    for (int c = 0; c < kCtxMaxQuads; ++c) {
        const int q = tid + c * kCtxThreads;
        if (q >= nq) continue;
        float4* Tq = reinterpret_cast<float4*>(p.T + (size_t)b * n_seg * D) + q;
        auto put = [&](int k, const float4& val) {                                // T[b,k] (+ its compact copy)
            Tq[(size_t)k * nq] = val;
            if (p.Tc != nullptr && s_segskip[k] == 0)
                reinterpret_cast<float4*>(p.Tc + (size_t)p.seg_pos[(size_t)b * n_seg + k] * D)[q] = val;
        };
        // empty segments are 0 -- except that with the compact output (Tc: the caller evaluates featureAffine and the pooling
        // on the visible segments only) a masked segment's row of T is never read by anyone, so its zeros are not written
        // (11 of 19 rows per user in MIND-shaped data: a fifth of this kernel's DRAM traffic)
        for (int k = 0; k < n_seg; ++k)
            if (s_start[k + 1] == s_start[k] && !(p.Tc != nullptr && s_segskip[k] != 0)) put(k, make_float4(0.f, 0.f, 0.f, 0.f));
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int cur = n_live > 0 ? s_seg[s_order[0]] : 0;
        for (int e0 = 0; e0 < n_live; e0 += kRows) {
            float4 x[kRows];
            float al[kRows];
            int sg[kRows];
#pragma unroll
            for (int u = 0; u < kRows; ++u) {
                const int t = s_order[min(e0 + u, n_live - 1)];
                sg[u] = s_seg[t];
                al[u] = s_alpha[t];
                x[u] = reinterpret_cast<const float4*>(Xh + (size_t)t * D)[q];
            }
#pragma unroll
            for (int u = 0; u < kRows; ++u) {
                if (e0 + u < n_live) {                                        // CTA-uniform control flow (shared-memory data only)
                    if (sg[u] != cur) {
                        put(cur, acc);
                        acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        cur = sg[u];
                    }
                    // alpha * x is rounded before the add in the reference (alpha * X then scatter_add): no fma here
                    acc.x = __fadd_rn(acc.x, __fmul_rn(al[u], x[u].x)); acc.y = __fadd_rn(acc.y, __fmul_rn(al[u], x[u].y));
                    acc.z = __fadd_rn(acc.z, __fmul_rn(al[u], x[u].z)); acc.w = __fadd_rn(acc.w, __fmul_rn(al[u], x[u].w));
                }
            }
        }
        if (n_live > 0) put(cur, acc);
    }
}

inline int launch_topic_segment_fwd(const float* Xu, int64_t strideX, const float* v, int ldv, const int64_t* cidx, float* T,
                                    float* alpha_out, int32_t* err_flag, const int32_t* src_index, const uint8_t* cmask,
                                    float* Tc, const int32_t* seg_pos, int B, int H, int n_seg, int D, cudaStream_t st) {
    if (B <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(Xu && v && cidx && T, "digat_topic_segment_fwd: null pointer");
    DIGAT_REQUIRE(H >= 1 && H <= kCtxMaxItems && n_seg >= 1 && n_seg <= kCtxMaxItems,
                  "digat_topic_segment_fwd: H=%d / n_seg=%d outside [1,%d]", H, n_seg, kCtxMaxItems);
    DIGAT_REQUIRE(D >= 4 && (D & 3) == 0 && D <= 4 * kCtxThreads * kCtxMaxQuads, "digat_topic_segment_fwd: bad D=%d", D);
    DIGAT_REQUIRE((strideX & 3) == 0 && aligned16(Xu) && aligned16(v) && aligned16(T),
                  "digat_topic_segment_fwd: pointers/strides must be 16-byte aligned");
    if (B <= 0) return DIGAT_OK;
    DIGAT_REQUIRE((ldv & 3) == 0 && ldv >= D, "digat_topic_segment_fwd: ldv must be a multiple of 4 and >= D");
    DIGAT_REQUIRE((Tc == nullptr) == (seg_pos == nullptr) && (Tc == nullptr || (cmask != nullptr && aligned16(Tc))),
                  "digat_topic_segment_fwd: Tc, seg_pos and cmask go together");
    SegArgs a{Xu, strideX, v, ldv, cidx, T, alpha_out, B, H, n_seg, D, err_flag, src_index, cmask, Tc, seg_pos};
    topic_segment_fwd_kernel<<<B, kCtxThreads, 0, st>>>(a);
    return check_launch("digat_topic_segment_fwd");
}

}  // namespace digat
