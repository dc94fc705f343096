// Fused Eq. (8) graph-attention layer, forward (replaces reference graphEncoders.py:150-153 / 170-173).
//
//   s_ij = sum_d a_d * relu((k3_d + K1_jd) + K2_id)      -- the [B,n,n,D] broadcast tensor is never materialised
//   e_ij = leaky_relu(s_ij, 0.2);  m_ij = adj_ij ? e_ij : -1e9;  alpha_i: = softmax_j(m_i:)
//   Y_i  = relu(sum_j alpha_ij h_j) + X_i
//
// One CTA owns R whole graphs (batch rows).  Three phases share one shared-memory arena:
//   phase 1  D is streamed in chunks: U = k3 + K1 and K2 chunks are staged in smem ([node][d], padded rows);
//            every thread owns a 4x4 set of (i,j) pairs (rows interleaved by nt so that consecutive lanes read
//            consecutive smem rows -> conflict-free LDS.128) and keeps the 16 partial dot products in registers;
//   phase 2  one warp per query node i: leaky-relu, mask, max / sum by warp shuffles, exp, normalise; the weights
//            stay in smem, transposed (St[j][i]) so that phase 3 reads 4 query rows per LDS.128;
//   phase 3  h is streamed in chunks; a thread owns 8 query rows x 4 features and accumulates alpha * h over j,
//            then applies relu + residual and writes Y with streaming 128-bit stores.
// HBM traffic per graph: read P (3nD) + X (nD) + adj (n^2 bytes) + k3, write Y (nD): 5nD*4 + n^2 bytes.
#pragma once
#include "common.cuh"

namespace digat {

constexpr int kPairThreads = 320;   // 10 warps: 289 pair tiles of a 68-node user graph fit in one pass
constexpr int kPairMaxNodes = 128;

struct PairAttnGeom {
    int R;        // graphs per CTA
    int nt;       // pair tiles per dimension = ceil(n/4)
    int dc1;      // phase-1 feature chunk (multiple of 4)
    int ld1;      // smem row stride of the phase-1 chunk (dc1 + 4 floats; odd multiple of 4 -> conflict-free)
    int dc3;      // phase-3 feature chunk
    int ld3;      // dc3 + 4
    int lds;      // leading dim of the transposed score matrix St[j][i] (multiple of 8, >= n)
    int arena;    // floats of the chunk arena per graph
    size_t smem;  // dynamic shared memory bytes
};

struct PairAttnArgs {
    const float* P; int ldp;
    const float* k3; const float* a; const uint8_t* adj; const float* X;
    float* Y; float* alpha_out;
    int B, n, D;
};

template <bool kSingleTile>
__global__ void __launch_bounds__(kPairThreads, 2)
graph_layer_fwd_kernel(PairAttnArgs p, PairAttnGeom g) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n, D = p.D, R = g.R;
    const int b0 = blockIdx.x * R;
    const int Rv = min(R, p.B - b0);                 // graphs actually present in this CTA

    float* a_s = smem;                                // [D]
    float* k3_s = a_s + D;                            // [R][D]
    float* St = k3_s + (size_t)R * D;                 // [R][n][lds]   St[r][j*lds + i]
    float* arena = St + (size_t)R * n * g.lds;        // [R][arena]

    for (int i = tid; i < D / 4; i += kPairThreads)
        reinterpret_cast<float4*>(a_s)[i] = reinterpret_cast<const float4*>(p.a)[i];
    for (int i = tid; i < Rv * (D / 4); i += kPairThreads) {
        int r = i / (D / 4), q = i % (D / 4);
        reinterpret_cast<float4*>(k3_s + (size_t)r * D)[q] =
            reinterpret_cast<const float4*>(p.k3 + (size_t)(b0 + r) * D)[q];
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 1: pair scores
    const int nt = g.nt, tiles_per_graph = nt * nt, tiles = Rv * tiles_per_graph;
    float acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;

    for (int c0 = 0; c0 < D; c0 += g.dc1) {
        const int w = min(g.dc1, D - c0), wq = w >> 2;
        // stage U = k3 + K1 (rounded exactly like the reference's first add) and K2
        for (int it = tid; it < Rv * n * wq; it += kPairThreads) {
            const int q = it % wq, node = (it / wq) % n, r = it / (wq * n);
            const float* prow = p.P + ((size_t)(b0 + r) * n + node) * p.ldp + c0 + 4 * q;
            float4 k1 = ldg_stream(reinterpret_cast<const float4*>(prow + D));
            float4 k2 = ldg_stream(reinterpret_cast<const float4*>(prow + 2 * D));
            const float4 kk = *reinterpret_cast<const float4*>(k3_s + (size_t)r * D + c0 + 4 * q);
            k1.x = kk.x + k1.x; k1.y = kk.y + k1.y; k1.z = kk.z + k1.z; k1.w = kk.w + k1.w;
            float* base = arena + (size_t)r * g.arena;
            *reinterpret_cast<float4*>(base + node * g.ld1 + 4 * q) = k1;                    // U  [n][ld1]
            *reinterpret_cast<float4*>(base + (n + node) * g.ld1 + 4 * q) = k2;              // K2 [n][ld1]
        }
        __syncthreads();
        for (int t = tid; t < tiles; t += kPairThreads) {
            const int r = t / tiles_per_graph, tt = t % tiles_per_graph;
            const int ti = tt / nt, tj = tt % nt;
            const float* Us = arena + (size_t)r * g.arena;
            const float* K2s = Us + n * g.ld1;
            int io[4], jo[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                io[x] = min(ti + nt * x, n - 1) * g.ld1;
                jo[x] = min(tj + nt * x, n - 1) * g.ld1;
            }
            if (!kSingleTile) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
            }
            const float4* a4 = reinterpret_cast<const float4*>(a_s + c0);
#pragma unroll 2
            for (int q = 0; q < wq; ++q) {
                const float4 av = a4[q];
                float4 k2[4], u[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    k2[x] = *reinterpret_cast<const float4*>(K2s + io[x] + 4 * q);
                    u[x] = *reinterpret_cast<const float4*>(Us + jo[x] + 4 * q);
                }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        float s = acc[x][y];
                        s = fmaf(av.x, fmaxf(u[y].x + k2[x].x, 0.f), s);
                        s = fmaf(av.y, fmaxf(u[y].y + k2[x].y, 0.f), s);
                        s = fmaf(av.z, fmaxf(u[y].z + k2[x].z, 0.f), s);
                        s = fmaf(av.w, fmaxf(u[y].w + k2[x].w, 0.f), s);
                        acc[x][y] = s;
                    }
            }
            if (!kSingleTile) {
                float* S = St + (size_t)r * n * g.lds;
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const int i = ti + nt * x, j = tj + nt * y;
                        if (i < n && j < n) {
                            float* dst = S + j * g.lds + i;
                            *dst = (c0 == 0 ? 0.f : *dst) + acc[x][y];
                        }
                    }
            }
        }
        __syncthreads();
    }
    if (kSingleTile && tid < tiles) {
        const int r = tid / tiles_per_graph, tt = tid % tiles_per_graph;
        const int ti = tt / nt, tj = tt % nt;
        float* S = St + (size_t)r * n * g.lds;
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const int i = ti + nt * x, j = tj + nt * y;
                if (i < n && j < n) S[j * g.lds + i] = acc[x][y];
            }
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 2: masked softmax per query node
    for (int row = warp; row < Rv * n; row += kPairThreads / 32) {
        const int r = row / n, i = row % n;
        float* S = St + (size_t)r * n * g.lds;
        const uint8_t* adj = p.adj + ((size_t)(b0 + r) * n + i) * n;
        float v[kPairMaxNodes / 32];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < kPairMaxNodes / 32; ++k) {
            const int j = lane + 32 * k;
            float m = -INFINITY;
            if (j < n) {
                const float s = S[j * g.lds + i];
                const float e = s > 0.f ? s : s * kLeakySlope;
                m = adj[j] != 0 ? e : kNegFill;
            }
            v[k] = m;
            mx = fmaxf(mx, m);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < kPairMaxNodes / 32; ++k) {
            const int j = lane + 32 * k;
            v[k] = (j < n) ? expf(v[k] - mx) : 0.f;
            sum += v[k];
        }
        sum = warp_sum(sum);
#pragma unroll
        for (int k = 0; k < kPairMaxNodes / 32; ++k) {
            const int j = lane + 32 * k;
            if (j < n) {
                const float al = v[k] / sum;
                S[j * g.lds + i] = al;
                if (p.alpha_out != nullptr) p.alpha_out[((size_t)(b0 + r) * n + i) * n + j] = al;
            }
        }
    }
    // padding columns i in [n, lds) are read (and discarded) by phase 3: keep them finite
    for (int it = tid; it < Rv * n * (g.lds - n); it += kPairThreads) {
        const int pad = g.lds - n;
        const int i = n + it % pad, j = (it / pad) % n, r = it / (pad * n);
        St[(size_t)r * n * g.lds + j * g.lds + i] = 0.f;
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 3: Y = relu(alpha * h) + X
    const int nit = (n + 7) >> 3;
    for (int c0 = 0; c0 < D; c0 += g.dc3) {
        const int w = min(g.dc3, D - c0), wq = w >> 2;
        for (int it = tid; it < Rv * n * wq; it += kPairThreads) {
            const int q = it % wq, node = (it / wq) % n, r = it / (wq * n);
            const float4 h = ldg_stream(reinterpret_cast<const float4*>(
                p.P + ((size_t)(b0 + r) * n + node) * p.ldp + c0 + 4 * q));
            *reinterpret_cast<float4*>(arena + (size_t)r * g.arena + node * g.ld3 + 4 * q) = h;
        }
        __syncthreads();
        const int items = Rv * nit * wq;
        for (int it = tid; it < items; it += kPairThreads) {
            const int q = it % wq, ib = (it / wq) % nit, r = it / (wq * nit);
            const int i0 = ib * 8;
            const float* Hs = arena + (size_t)r * g.arena + 4 * q;
            const float* S = St + (size_t)r * n * g.lds + i0;
            float4 o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
            for (int j = 0; j < n; ++j) {
                const float4 h = *reinterpret_cast<const float4*>(Hs + j * g.ld3);
                const float4 al0 = *reinterpret_cast<const float4*>(S + j * g.lds);
                const float4 al1 = *reinterpret_cast<const float4*>(S + j * g.lds + 4);
                const float al[8] = {al0.x, al0.y, al0.z, al0.w, al1.x, al1.y, al1.z, al1.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    o[k].x = fmaf(al[k], h.x, o[k].x);
                    o[k].y = fmaf(al[k], h.y, o[k].y);
                    o[k].z = fmaf(al[k], h.z, o[k].z);
                    o[k].w = fmaf(al[k], h.w, o[k].w);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = i0 + k;
                if (i < n) {
                    const size_t off = ((size_t)(b0 + r) * n + i) * D + c0 + 4 * q;
                    const float4 x = ldg_stream(reinterpret_cast<const float4*>(p.X + off));
                    float4 y;
                    y.x = fmaxf(o[k].x, 0.f) + x.x;
                    y.y = fmaxf(o[k].y, 0.f) + x.y;
                    y.z = fmaxf(o[k].z, 0.f) + x.z;
                    y.w = fmaxf(o[k].w, 0.f) + x.w;
                    stg_stream(reinterpret_cast<float4*>(p.Y + off), y);
                }
            }
        }
        __syncthreads();
    }
}

// Host-side geometry: graphs per CTA and chunk widths for (n, D) under a shared-memory budget.
inline bool pair_attn_geometry(int n, int D, int B, PairAttnGeom* g) {
    const int nt = (n + 3) / 4;
    const int tiles = nt * nt;
    int dc1 = 80;
    if (dc1 > D) dc1 = D;
    const int ld1 = dc1 + 4;
    const int lds = ((n + 7) / 8) * 8 + 4;        // multiple of 4; +4 shifts consecutive j by 4 banks
    const int arena = 2 * n * ld1;
    const size_t per_graph = (size_t)(D + n * lds + arena) * sizeof(float);
    const size_t budget = 72 * 1024;               // keeps 3 CTAs per SM for the 68-node user graph
    int R = kPairThreads / tiles;
    if (R < 1) R = 1;
    while (R > 1 && (size_t)D * 4 + R * per_graph > budget) --R;
    if (R > B) R = B > 0 ? B : 1;
    // phase-3 chunk: widest multiple of 4 that fits the arena and maximises lane utilisation
    const int nit = (n + 7) / 8;
    const int max_dc3 = (arena / n) - 4;
    int best = 4; double best_eff = -1.0;
    for (int dc3 = 16; dc3 <= max_dc3 && dc3 <= D; dc3 += 4) {
        long busy = 0, slots = 0;
        for (int c0 = 0; c0 < D; c0 += dc3) {
            const int w = (D - c0 < dc3) ? D - c0 : dc3;
            const long items = (long)R * nit * (w / 4);
            busy += items;
            slots += ((items + kPairThreads - 1) / kPairThreads) * kPairThreads + 64;   // +64: per-chunk sync cost
        }
        const double eff = (double)busy / (double)slots;
        if (eff > best_eff) { best_eff = eff; best = dc3; }
    }
    g->R = R; g->nt = nt; g->dc1 = dc1; g->ld1 = ld1; g->dc3 = best; g->ld3 = best + 4; g->lds = lds;
    g->arena = arena;
    g->smem = (size_t)D * 4 + (size_t)R * per_graph;
    return true;
}

inline int launch_graph_layer_fwd(const float* P, int ldp, const float* k3, const float* a, const uint8_t* adj,
                                  const float* X, float* Y, float* alpha_out, int B, int n, int D, cudaStream_t st) {
    DIGAT_REQUIRE(P && k3 && a && adj && X && Y, "digat_graph_layer_fwd: null pointer");
    DIGAT_REQUIRE(B >= 0 && n >= 1 && n <= kPairMaxNodes, "digat_graph_layer_fwd: n=%d outside [1,%d]", n, kPairMaxNodes);
    DIGAT_REQUIRE(D >= 4 && (D & 3) == 0 && D <= 1024, "digat_graph_layer_fwd: D=%d must be a multiple of 4 in [4,1024]", D);
    DIGAT_REQUIRE((ldp & 3) == 0 && ldp >= 3 * D, "digat_graph_layer_fwd: ldp=%d must be a multiple of 4 and >= 3D", ldp);
    DIGAT_REQUIRE(aligned16(P) && aligned16(k3) && aligned16(a) && aligned16(X) && aligned16(Y),
                  "digat_graph_layer_fwd: pointers must be 16-byte aligned");
    if (B == 0) return DIGAT_OK;
    PairAttnGeom g;
    pair_attn_geometry(n, D, B, &g);
    const DeviceInfo* di = device_info();
    if (!di) return fail(DIGAT_E_CUDA, "digat_graph_layer_fwd: no CUDA device");
    DIGAT_REQUIRE(g.smem <= (size_t)di->max_smem_optin, "digat_graph_layer_fwd: needs %zu B shared memory", g.smem);
    PairAttnArgs args{P, ldp, k3, a, adj, X, Y, alpha_out, B, n, D};
    const int grid = (B + g.R - 1) / g.R;
    const bool single = g.R * g.nt * g.nt <= kPairThreads;
    if (single) {
        DIGAT_CUDA(cudaFuncSetAttribute(graph_layer_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        graph_layer_fwd_kernel<true><<<grid, kPairThreads, g.smem, st>>>(args, g);
    } else {
        DIGAT_CUDA(cudaFuncSetAttribute(graph_layer_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        graph_layer_fwd_kernel<false><<<grid, kPairThreads, g.smem, st>>>(args, g);
    }
    return check_launch("digat_graph_layer_fwd");
}

}  // namespace digat
