// News (title) encoder, MSA variant -- reference newsEncoders.py:58-82 + layers.py:50-115.
//
//   w  = word_embedding[title]                                  [titles, T, E]      (digat_gather_rows_i32)
//   QKV = w [W_Q; W_K; W_V]^T + [b_Q; 0; b_V]                   [titles*T, 3*h*dk]   (projection GEMM, one launch)
//   per head: A = Q K^T / sqrt(dk);  alpha = softmax_j(A);  O = alpha V;  H = relu(concat heads)      msa_attention_kernel
//   att = tanh(H affine1^T + b1);  a_t = att_t . w2;  alpha = softmax_t(mask(a));  out = sum_t alpha_t H_t   (GEMM +) additive_pool_kernel
//
// msa_attention_kernel: a title has T <= 32 tokens, so ONE WARP owns one (title, head): lane i is query token i.  K_h and
// V_h (T x dk floats each) sit in the warp's slice of shared memory and are read as broadcasts; the lane keeps its q row,
// its T scores and its dk outputs in registers -- the softmax needs no shuffle at all.  The self-attention has no padding
// mask (layers.py:88-97 applies none); padded tokens are masked only by the pooling afterwards.
// HBM traffic per title: read T*3*h*dk*4, write T*h*dk*4 bytes (205 KB at T=32, h*dk=400): the kernel is HBM-bound.
//
// additive_pool_kernel: one CTA per title, a warp per token for the tanh / dot pass, then the masked softmax over the T
// tokens and the weighted sum of the H rows (each read once from L2/DRAM).
#pragma once
#include "common.cuh"
#include "tma.cuh"   // ensure_dynamic_smem

namespace digat {

constexpr int kMsaMaxT = 32;       // tokens per title (config.max_title_length, default 32)
constexpr int kMsaMaxDk = 32;      // head dimension (config.MSA_head_dim, default 25)
constexpr int kMsaWarps = 8;       // (title, head) pairs per CTA

__global__ void __launch_bounds__(kMsaWarps * 32)
msa_attention_kernel(const float* __restrict__ QKV, int ld, float* __restrict__ H, int ldh, int64_t n_titles, int T, int heads,
                     int dk, float inv_scale_div) {
    extern __shared__ __align__(16) float msa_smem[];               // [kMsaWarps][2][T * dkp], rows padded to float4s (zeros)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t unit = (int64_t)blockIdx.x * kMsaWarps + warp;    // (title, head)
    if (unit >= n_titles * heads) return;
    const int64_t title = unit / heads;
    const int head = (int)(unit - title * heads);
    const int hd = heads * dk;
    const int dkp = (dk + 3) & ~3, nq = dkp >> 2;                   // every broadcast read of a K / V row is an LDS.128
    float* Ks = msa_smem + (size_t)warp * 2 * T * dkp;
    float* Vs = Ks + T * dkp;
    const float* base = QKV + (size_t)title * T * ld + head * dk;   // Q block; K at +hd, V at +2*hd
    for (int e = lane; e < T * dkp; e += 32) {
        const int t = e / dkp, d = e - t * dkp;
        Ks[e] = d < dk ? base[(size_t)t * ld + hd + d] : 0.f;
        Vs[e] = d < dk ? base[(size_t)t * ld + 2 * hd + d] : 0.f;
    }
    float q[kMsaMaxDk];
#pragma unroll
    for (int d = 0; d < kMsaMaxDk; ++d) q[d] = (lane < T && d < dk) ? base[(size_t)lane * ld + d] : 0.f;
    __syncwarp();
    float s[kMsaMaxT];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kMsaMaxT; ++j) {
        float acc = 0.f;
        if (j < T) {
            const float4* kr = reinterpret_cast<const float4*>(Ks + j * dkp);
#pragma unroll
            for (int d4 = 0; d4 < kMsaMaxDk / 4; ++d4)
                if (d4 < nq) {                                       // (same ascending-d fma chain as a scalar loop: q is 0 past dk)
                    const float4 k = kr[d4];
                    acc = fmaf(q[4 * d4 + 0], k.x, acc); acc = fmaf(q[4 * d4 + 1], k.y, acc);
                    acc = fmaf(q[4 * d4 + 2], k.z, acc); acc = fmaf(q[4 * d4 + 3], k.w, acc);
                }
            acc = acc / inv_scale_div;                               // the reference divides by sqrt(dk) (layers.py:89)
            mx = fmaxf(mx, acc);
        }
        s[j] = acc;
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kMsaMaxT; ++j) {
        s[j] = j < T ? expf(s[j] - mx) : 0.f;
        sum += s[j];
    }
    float o[kMsaMaxDk];
#pragma unroll
    for (int d = 0; d < kMsaMaxDk; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < kMsaMaxT; ++j) {
        if (j < T) {
            const float p = s[j] / sum;
            const float4* vr = reinterpret_cast<const float4*>(Vs + j * dkp);
#pragma unroll
            for (int d4 = 0; d4 < kMsaMaxDk / 4; ++d4)
                if (d4 < nq) {
                    const float4 v = vr[d4];
                    o[4 * d4 + 0] = fmaf(p, v.x, o[4 * d4 + 0]); o[4 * d4 + 1] = fmaf(p, v.y, o[4 * d4 + 1]);
                    o[4 * d4 + 2] = fmaf(p, v.z, o[4 * d4 + 2]); o[4 * d4 + 3] = fmaf(p, v.w, o[4 * d4 + 3]);
                }
        }
    }
    if (lane < T) {
        float* out = H + ((size_t)title * T + lane) * ldh + head * dk;
#pragma unroll
        for (int d = 0; d < kMsaMaxDk; ++d)
            if (d < dk) out[d] = fmaxf(o[d], 0.f);                   // F.relu (newsEncoders.py:78)
    }
}

inline int launch_msa_attention(const float* QKV, int ld, float* H, int ldh, int64_t n_titles, int T, int heads, int dk,
                                cudaStream_t st) {
    if (n_titles <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(QKV && H, "digat_msa_attention_fwd: null pointer");
    DIGAT_REQUIRE(T >= 1 && T <= kMsaMaxT && dk >= 1 && dk <= kMsaMaxDk && heads >= 1,
                  "digat_msa_attention_fwd: needs max_title_length <= %d and head_dim <= %d (T=%d, dk=%d)", kMsaMaxT, kMsaMaxDk, T, dk);
    DIGAT_REQUIRE(ld >= 3 * heads * dk && ldh >= heads * dk, "digat_msa_attention_fwd: leading dimension too small");
    const int64_t units = n_titles * heads;
    const size_t smem = (size_t)kMsaWarps * 2 * T * ((dk + 3) & ~3) * sizeof(float);
    DIGAT_REQUIRE(units / kMsaWarps + 1 < (1LL << 31), "digat_msa_attention_fwd: too many titles for one launch");
    if (int rc_ = ensure_dynamic_smem(msa_attention_kernel, smem)) return rc_;
    msa_attention_kernel<<<(unsigned)((units + kMsaWarps - 1) / kMsaWarps), kMsaWarps * 32, smem, st>>>(
        QKV, ld, H, ldh, n_titles, T, heads, dk, sqrtf((float)dk));
    return check_launch("digat_msa_attention_fwd");
}

// out[title] = sum_t softmax_t(mask(att_t . w2)) H[title, t, :],  att = tanh(pre-activation rows of the affine1 GEMM)
constexpr int kPoolThreads = 128;

__global__ void __launch_bounds__(kPoolThreads)
additive_pool_kernel(const float* __restrict__ att_pre, int lda, const float* __restrict__ w2, const float* __restrict__ H,
                     int ldh, const uint8_t* __restrict__ mask, float* __restrict__ out, int ldo, int T, int A, int D) {
    __shared__ float a_s[kMsaMaxT];
    const int64_t title = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = warp; t < T; t += kPoolThreads / 32) {
        const float* row = att_pre + ((size_t)title * T + t) * lda;
        float acc = 0.f;
        for (int c = lane; c < A; c += 32) acc = fmaf(tanhf(row[c]), w2[c], acc);
        acc = warp_sum(acc);
        if (lane == 0) a_s[t] = mask[(size_t)title * T + t] != 0 ? acc : kNegFill;      // masked_fill(mask == 0, -1e9), layers.py:111
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int t = 0; t < T; ++t) mx = fmaxf(mx, a_s[t]);
    float sum = 0.f;
    for (int t = 0; t < T; ++t) sum += expf(a_s[t] - mx);
    for (int q = threadIdx.x; q < D / 4; q += kPoolThreads) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t < T; ++t) {
            const float p = expf(a_s[t] - mx) / sum;
            const float4 h = reinterpret_cast<const float4*>(H + ((size_t)title * T + t) * ldh)[q];
            acc.x = fmaf(p, h.x, acc.x); acc.y = fmaf(p, h.y, acc.y); acc.z = fmaf(p, h.z, acc.z); acc.w = fmaf(p, h.w, acc.w);
        }
        reinterpret_cast<float4*>(out + (size_t)title * ldo)[q] = acc;
    }
}

inline int launch_additive_pool(const float* att_pre, int lda, const float* w2, const float* H, int ldh, const uint8_t* mask,
                                float* out, int ldo, int64_t n_titles, int T, int A, int D, cudaStream_t st) {
    if (n_titles <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(att_pre && w2 && H && mask && out, "digat_additive_pool_fwd: null pointer");
    DIGAT_REQUIRE(T >= 1 && T <= kMsaMaxT && A >= 1 && D >= 4 && (D & 3) == 0 && (ldh & 3) == 0 && (ldo & 3) == 0 &&
                  lda >= A && ldh >= D && ldo >= D, "digat_additive_pool_fwd: bad sizes (T=%d, A=%d, D=%d)", T, A, D);
    DIGAT_REQUIRE(aligned16(H) && aligned16(out), "digat_additive_pool_fwd: H / out must be 16-byte aligned");
    DIGAT_REQUIRE(n_titles < (1LL << 31), "digat_additive_pool_fwd: too many titles for one launch");
    additive_pool_kernel<<<(unsigned)n_titles, kPoolThreads, 0, st>>>(att_pre, lda, w2, H, ldh, mask, out, ldo, T, A, D);
    return check_launch("digat_additive_pool_fwd");
}

// ------------------------------------------------------------------------------------------------ backward (training)
// msa_attention_bwd_kernel: one warp per (title, head), as the forward.  The T x T attention matrix is recomputed from the
// saved Q | K | V rows; lane i owns query row i (p_i., d p_i., dS_i. in registers -> dQ_i), writes its rows of P and dS to the
// warp's shared-memory slice, and then owns key row i for the transposed sums dK_i = sum_q dS[q][i] Q_q, dV_i = sum_q P[q][i] dO_q.
// dO = dH * 1[H > 0] (the relu of newsEncoders.py:78) is applied while staging.
constexpr int kMsaBwdWarps = 4;

__global__ void __launch_bounds__(kMsaBwdWarps * 32)
msa_attention_bwd_kernel(const float* __restrict__ QKV, int ld, const float* __restrict__ H, int ldh, const float* __restrict__ dH,
                         int lddh, float* __restrict__ dQKV, int ldd, int64_t n_titles, int T, int heads, int dk, float scale_div) {
    extern __shared__ __align__(16) float msa_smem[];               // per warp: Q, K, V, dO [T][dkp] each, P, dS [T][33] each
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t unit = (int64_t)blockIdx.x * kMsaBwdWarps + warp;
    if (unit >= n_titles * heads) return;
    const int64_t title = unit / heads;
    const int head = (int)(unit - title * heads);
    const int hd = heads * dk;
    const int dkp = (dk + 3) & ~3, nq = dkp >> 2;                   // rows padded to float4s (zeros): every broadcast read is an LDS.128
    float* Qs = msa_smem + (size_t)warp * ((4 * T * dkp + 2 * T * 33 + 3) & ~3);       // 16-byte aligned slices
    float* Ks = Qs + T * dkp;
    float* Vs = Ks + T * dkp;
    float* dOs = Vs + T * dkp;
    float* Ps = dOs + T * dkp;
    float* dSs = Ps + T * 33;
    const float* base = QKV + (size_t)title * T * ld + head * dk;
    for (int e = lane; e < T * dkp; e += 32) {
        const int t = e / dkp, d = e - t * dkp;
        const bool in = d < dk;
        const size_t hrow = (size_t)title * T + t;
        Qs[e] = in ? base[(size_t)t * ld + d] : 0.f;
        Ks[e] = in ? base[(size_t)t * ld + hd + d] : 0.f;
        Vs[e] = in ? base[(size_t)t * ld + 2 * hd + d] : 0.f;
        dOs[e] = (in && H[hrow * ldh + head * dk + d] > 0.f) ? dH[hrow * lddh + head * dk + d] : 0.f;
    }
    __syncwarp();
    const int i = lane;
    if (i < T) {
        float4 q[kMsaMaxDk / 4], go[kMsaMaxDk / 4], dq[kMsaMaxDk / 4];
#pragma unroll
        for (int d4 = 0; d4 < kMsaMaxDk / 4; ++d4) {
            q[d4] = d4 < nq ? reinterpret_cast<const float4*>(Qs + i * dkp)[d4] : make_float4(0.f, 0.f, 0.f, 0.f);
            go[d4] = d4 < nq ? reinterpret_cast<const float4*>(dOs + i * dkp)[d4] : make_float4(0.f, 0.f, 0.f, 0.f);
            dq[d4] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float pr[kMsaMaxT], dp[kMsaMaxT];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kMsaMaxT; ++j) {
            float acc = 0.f, g = 0.f;
            if (j < T) {
                const float4* kr = reinterpret_cast<const float4*>(Ks + j * dkp);
                const float4* vr = reinterpret_cast<const float4*>(Vs + j * dkp);
#pragma unroll
                for (int d4 = 0; d4 < kMsaMaxDk / 4; ++d4)
                    if (d4 < nq) {
                        const float4 k = kr[d4], v = vr[d4];
                        acc = fmaf(q[d4].x, k.x, acc); acc = fmaf(q[d4].y, k.y, acc); acc = fmaf(q[d4].z, k.z, acc); acc = fmaf(q[d4].w, k.w, acc);
                        g = fmaf(go[d4].x, v.x, g); g = fmaf(go[d4].y, v.y, g); g = fmaf(go[d4].z, v.z, g); g = fmaf(go[d4].w, v.w, g);
                    }
                acc = acc / scale_div;
                mx = fmaxf(mx, acc);
            }
            pr[j] = acc;
            dp[j] = g;
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < kMsaMaxT; ++j) {
            pr[j] = j < T ? expf(pr[j] - mx) : 0.f;
            sum += pr[j];
        }
        float delta = 0.f;
#pragma unroll
        for (int j = 0; j < kMsaMaxT; ++j) {
            pr[j] = pr[j] / sum;
            delta = fmaf(pr[j], dp[j], delta);
        }
#pragma unroll
        for (int j = 0; j < kMsaMaxT; ++j) {
            if (j < T) {
                const float ds = pr[j] * (dp[j] - delta) / scale_div;
                Ps[i * 33 + j] = pr[j];
                dSs[i * 33 + j] = ds;
                const float4* kr = reinterpret_cast<const float4*>(Ks + j * dkp);
#pragma unroll
                for (int d4 = 0; d4 < kMsaMaxDk / 4; ++d4)
                    if (d4 < nq) {
                        const float4 k = kr[d4];
                        dq[d4].x = fmaf(ds, k.x, dq[d4].x); dq[d4].y = fmaf(ds, k.y, dq[d4].y);
                        dq[d4].z = fmaf(ds, k.z, dq[d4].z); dq[d4].w = fmaf(ds, k.w, dq[d4].w);
                    }
            }
        }
        float* out = dQKV + ((size_t)title * T + i) * ldd + head * dk;
#pragma unroll
        for (int d4 = 0; d4 < kMsaMaxDk / 4; ++d4) {
            const float v4[4] = {dq[d4].x, dq[d4].y, dq[d4].z, dq[d4].w};
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (4 * d4 + c < dk) out[4 * d4 + c] = v4[c];
        }
    }
    __syncwarp();
    if (i < T) {                                                    // lane i now owns KEY / VALUE row i
        float4 gk[kMsaMaxDk / 4], gv[kMsaMaxDk / 4];
#pragma unroll
        for (int d4 = 0; d4 < kMsaMaxDk / 4; ++d4) gk[d4] = gv[d4] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int qr = 0; qr < T; ++qr) {
            const float ds = dSs[qr * 33 + i], p = Ps[qr * 33 + i];
            const float4* qrow = reinterpret_cast<const float4*>(Qs + qr * dkp);
            const float4* grow = reinterpret_cast<const float4*>(dOs + qr * dkp);
#pragma unroll
            for (int d4 = 0; d4 < kMsaMaxDk / 4; ++d4)
                if (d4 < nq) {
                    const float4 qq = qrow[d4], gg = grow[d4];
                    gk[d4].x = fmaf(ds, qq.x, gk[d4].x); gk[d4].y = fmaf(ds, qq.y, gk[d4].y);
                    gk[d4].z = fmaf(ds, qq.z, gk[d4].z); gk[d4].w = fmaf(ds, qq.w, gk[d4].w);
                    gv[d4].x = fmaf(p, gg.x, gv[d4].x); gv[d4].y = fmaf(p, gg.y, gv[d4].y);
                    gv[d4].z = fmaf(p, gg.z, gv[d4].z); gv[d4].w = fmaf(p, gg.w, gv[d4].w);
                }
        }
        float* out = dQKV + ((size_t)title * T + i) * ldd + head * dk;
#pragma unroll
        for (int d4 = 0; d4 < kMsaMaxDk / 4; ++d4) {
            const float k4[4] = {gk[d4].x, gk[d4].y, gk[d4].z, gk[d4].w};
            const float v4[4] = {gv[d4].x, gv[d4].y, gv[d4].z, gv[d4].w};
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (4 * d4 + c < dk) {
                    out[hd + 4 * d4 + c] = k4[c];
                    out[2 * hd + 4 * d4 + c] = v4[c];
                }
        }
    }
}

inline int launch_msa_attention_bwd(const float* QKV, int ld, const float* H, int ldh, const float* dH, int lddh, float* dQKV,
                                    int ldd, int64_t n_titles, int T, int heads, int dk, cudaStream_t st) {
    if (n_titles <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(QKV && H && dH && dQKV, "digat_msa_attention_bwd: null pointer");
    DIGAT_REQUIRE(T >= 1 && T <= kMsaMaxT && dk >= 1 && dk <= kMsaMaxDk && heads >= 1,
                  "digat_msa_attention_bwd: needs max_title_length <= %d and head_dim <= %d (T=%d, dk=%d)", kMsaMaxT, kMsaMaxDk, T, dk);
    DIGAT_REQUIRE(ld >= 3 * heads * dk && ldd >= 3 * heads * dk && ldh >= heads * dk && lddh >= heads * dk,
                  "digat_msa_attention_bwd: leading dimension too small");
    const int64_t units = n_titles * heads;
    const int dkp = (dk + 3) & ~3;
    const size_t smem = (size_t)kMsaBwdWarps * ((4 * T * dkp + 2 * T * 33 + 3) & ~3) * sizeof(float);
    DIGAT_REQUIRE(units / kMsaBwdWarps + 1 < (1LL << 31), "digat_msa_attention_bwd: too many titles for one launch");
    if (int rc_ = ensure_dynamic_smem(msa_attention_bwd_kernel, smem)) return rc_;
    msa_attention_bwd_kernel<<<(unsigned)((units + kMsaBwdWarps - 1) / kMsaBwdWarps), kMsaBwdWarps * 32, smem, st>>>(
        QKV, ld, H, ldh, dH, lddh, dQKV, ldd, n_titles, T, heads, dk, sqrtf((float)dk));
    return check_launch("digat_msa_attention_bwd");
}

// Backward of additive_pool_kernel (layers.py:107-115).  One CTA per title:
//   alpha recomputed;  dalpha_t = dout . H_t;  da_t = alpha_t (dalpha_t - sum_k alpha_k dalpha_k)   (0 on masked tokens)
//   dH_t (through the weighted sum) = alpha_t dout;  datt_pre[t, c] = da_t w2[c] (1 - tanh^2);  dw2_part[title, c] = sum_t da_t tanh
// (the gradient through the affine1 GEMM reaches H via the caller's autograd and is added to dH there).
__global__ void __launch_bounds__(kPoolThreads)
additive_pool_bwd_kernel(const float* __restrict__ att_pre, int lda, const float* __restrict__ w2, const float* __restrict__ H,
                         int ldh, const uint8_t* __restrict__ mask, const float* __restrict__ dout, int ldo,
                         float* __restrict__ dH, int lddh, float* __restrict__ datt, int ldda, float* __restrict__ dw2_part,
                         int T, int A, int D) {
    __shared__ float a_s[kMsaMaxT], al_s[kMsaMaxT], ds_s[kMsaMaxT];
    __shared__ double da_s[kMsaMaxT];      // dalpha_t in double: alpha_t (dalpha_t - sum alpha dalpha) cancels (cf. topic_segment_bwd)
    const int64_t title = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* go = dout + (size_t)title * ldo;
    for (int t = warp; t < T; t += kPoolThreads / 32) {
        const float* row = att_pre + ((size_t)title * T + t) * lda;
        const float* hrow = H + ((size_t)title * T + t) * ldh;
        float acc = 0.f;
        double g = 0.0;
        for (int c = lane; c < A; c += 32) acc = fmaf(tanhf(row[c]), w2[c], acc);
        for (int d = lane; d < D; d += 32) g += (double)go[d] * hrow[d];
        acc = warp_sum(acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
        if (lane == 0) {
            a_s[t] = mask[(size_t)title * T + t] != 0 ? acc : kNegFill;
            da_s[t] = g;
        }
    }
    __syncthreads();
    if (threadIdx.x < T) {
        float mx = -INFINITY;
        for (int t = 0; t < T; ++t) mx = fmaxf(mx, a_s[t]);
        float sum = 0.f;
        for (int t = 0; t < T; ++t) sum += expf(a_s[t] - mx);
        al_s[threadIdx.x] = expf(a_s[threadIdx.x] - mx) / sum;
    }
    __syncthreads();
    if (threadIdx.x < T) {
        double tt = 0.0;
        for (int t = 0; t < T; ++t) tt += (double)al_s[t] * da_s[t];
        // a masked token's score is the constant fill value: no gradient (masked_fill), whatever alpha it ends up with
        ds_s[threadIdx.x] = mask[(size_t)title * T + threadIdx.x] != 0 ? (float)((double)al_s[threadIdx.x] * (da_s[threadIdx.x] - tt)) : 0.f;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < D / 4; q += kPoolThreads) {
        const float4 g = reinterpret_cast<const float4*>(go)[q];
        for (int t = 0; t < T; ++t) {
            const float p = al_s[t];
            reinterpret_cast<float4*>(dH + ((size_t)title * T + t) * lddh)[q] = make_float4(p * g.x, p * g.y, p * g.z, p * g.w);
        }
    }
    for (int c = threadIdx.x; c < A; c += kPoolThreads) {
        const float w = w2[c];
        float acc = 0.f;
        for (int t = 0; t < T; ++t) {
            const float th = tanhf(att_pre[((size_t)title * T + t) * lda + c]);
            const float ds = ds_s[t];
            datt[((size_t)title * T + t) * ldda + c] = ds * w * (1.f - th * th);
            acc = fmaf(ds, th, acc);
        }
        dw2_part[(size_t)title * A + c] = acc;
    }
}

inline int launch_additive_pool_bwd(const float* att_pre, int lda, const float* w2, const float* H, int ldh, const uint8_t* mask,
                                    const float* dout, int ldo, float* dH, int lddh, float* datt, int ldda, float* dw2_part,
                                    int64_t n_titles, int T, int A, int D, cudaStream_t st) {
    if (n_titles <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(att_pre && w2 && H && mask && dout && dH && datt && dw2_part, "digat_additive_pool_bwd: null pointer");
    DIGAT_REQUIRE(T >= 1 && T <= kMsaMaxT && A >= 1 && D >= 4 && (D & 3) == 0 && (ldh & 3) == 0 && (ldo & 3) == 0 && (lddh & 3) == 0 &&
                  lda >= A && ldda >= A && ldh >= D && ldo >= D && lddh >= D, "digat_additive_pool_bwd: bad sizes (T=%d, A=%d, D=%d)", T, A, D);
    DIGAT_REQUIRE(aligned16(dout) && aligned16(dH), "digat_additive_pool_bwd: dout / dH must be 16-byte aligned");
    DIGAT_REQUIRE(n_titles < (1LL << 31), "digat_additive_pool_bwd: too many titles for one launch");
    additive_pool_bwd_kernel<<<(unsigned)n_titles, kPoolThreads, 0, st>>>(att_pre, lda, w2, H, ldh, mask, dout, ldo, dH, lddh, datt,
                                                                          ldda, dw2_part, T, A, D);
    return check_launch("digat_additive_pool_bwd");
}

// dtable[idx[r], :] += src[r, :]   (backward of the embedding gather; float atomics, like torch's embedding backward)
__global__ void scatter_add_rows_kernel(float* __restrict__ dtable, int64_t n_table, const int32_t* __restrict__ idx,
                                        const float* __restrict__ src, int64_t lds, int64_t rows, int D) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * D) return;
    const int64_t r = i / D;
    const int d = (int)(i - r * D);
    const int64_t t = idx[r];
    if (t < 0 || t >= n_table) return;                              // (the forward gather reports out-of-range ids)
    atomicAdd(dtable + t * D + d, src[r * lds + d]);
}

// the same with one 16-byte reduction per quad (red.global.add.v4.f32, sm_90+): D, lds multiples of 4, 16-byte aligned bases
__global__ void scatter_add_rows_v4_kernel(float* __restrict__ dtable, int64_t n_table, const int32_t* __restrict__ idx,
                                           const float* __restrict__ src, int64_t lds, int64_t rows, int Dq) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * Dq) return;
    const int64_t r = i / Dq;
    const int q = (int)(i - r * Dq);
    const int64_t t = idx[r];
    if (t < 0 || t >= n_table) return;
    const float4 v = *reinterpret_cast<const float4*>(src + r * lds + 4 * q);
    atomicAdd(reinterpret_cast<float4*>(dtable + t * (int64_t)(4 * Dq)) + q, v);
}

inline int launch_scatter_add_rows(float* dtable, int64_t n_table, const int32_t* idx, const float* src, int64_t lds, int64_t rows,
                                   int D, cudaStream_t st) {
    if (rows <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(dtable && idx && src && D >= 1 && lds >= D, "digat_scatter_add_rows: null pointer or bad sizes");
    const int64_t total = rows * D;
    DIGAT_REQUIRE((total + 255) / 256 < (1LL << 31), "digat_scatter_add_rows: too many elements for one launch");
    if ((D & 3) == 0 && (lds & 3) == 0 && aligned16(dtable) && aligned16(src)) {
        const int64_t quads = rows * (D / 4);
        scatter_add_rows_v4_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, st>>>(dtable, n_table, idx, src, lds, rows, D / 4);
        return check_launch("digat_scatter_add_rows");
    }
    scatter_add_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dtable, n_table, idx, src, lds, rows, D);
    return check_launch("digat_scatter_add_rows");
}

}  // namespace digat
