// Fused Eq. (8) graph-attention layer, forward (replaces reference graphEncoders.py:150-153 / 170-173).
//
//   s_ij = sum_d a_d * relu(U_jd + K2_id),  U = k3 + K1   -- the [B,n,n,D] broadcast tensor is never materialised
//   e_ij = leaky_relu(s_ij, 0.2);  m_ij = adj_ij ? e_ij : -1e9;  alpha_i: = softmax_j(m_i:)
//   Y_i  = relu(sum_j alpha_ij h_j) + X_i
//
// U = fl(k3 + K1) is produced by the projection GEMM's epilogue (row-group bias), rounded exactly like the
// reference's first broadcast add, so this kernel only streams tiles of P = [h | U | K2].
//
// One CTA owns R whole graphs (batch rows).  The feature dimension is streamed through a 2-deep TMA pipeline
// (cp.async.bulk.tensor.2d -> mbarrier); all three phases share the two shared-memory buffers:
//   phase 1  per chunk of `dc` features: U and K2 tiles [R*n][dc] (dense rows; dc/4 odd -> consecutive rows start
//            4 banks apart, so the row-interleaved LDS.128 pattern is conflict-free).  Every thread owns a 4x4 set
//            of (i,j) pairs and keeps 16 x 2 partial dot products in registers; the inner loop is packed fp32x2
//            (FADD2 / FFMA2) with scalar FMNMX for the relu: 2 issue slots per (i,j,d) instead of 3;
//   phase 2  one warp per query node: leaky-relu, mask, max / sum by warp shuffles, exp, normalise; the weights stay
//            in smem, transposed (St[j][i]) so that phase 3 reads 8 query rows with two LDS.128;
//   phase 3  h streamed in chunks of 2*dc features; a thread owns 8 query rows x 4 features, accumulates
//            alpha * h over j (FFMA2), applies relu + residual and writes Y with streaming 128-bit stores.
// The loads of the next chunk (and of the first h chunks during phase 2) overlap the math of the current one.
// HBM traffic per graph: read P (3nD) + X (nD) + adj (n^2 bytes), write Y (nD): 5nD*4 + n^2 bytes.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace digat {

constexpr int kPairThreads = 320;   // 10 warps: the 289 pair tiles of a 68-node user graph fit in one pass
constexpr int kPairMaxNodes = 128;

struct PairAttnGeom {
    int R;           // graphs per CTA
    int nt;          // pair tiles per dimension = ceil(n/4)
    int dc;          // phase-1 feature chunk (multiple of 4; dc/4 odd when possible)
    int nch1;        // phase-1 chunks = ceil(D/dc)
    int dc3;         // phase-3 feature chunk = 2*dc
    int nch3;        // phase-3 chunks = ceil(D/dc3)
    int lds;         // leading dim of the transposed score matrix St[j][i] (multiple of 4, >= 8*ceil(n/8))
    int tile_floats; // floats of one [R*n][dc] tile, rounded up to 128 bytes
    size_t smem;     // dynamic shared memory bytes
};

struct PairAttnArgs {
    const float* P; int ldp;
    const float* a; const uint8_t* adj; const float* X;
    float* Y;
    int B, n, D;
    // training extras (all optional): dropout on the attention weights (reference graphEncoders.py:152/172) with a
    // caller-provided keep mask, and the tensors the backward pass needs
    const uint8_t* drop_keep;   // [B,n,n] 1 = keep; alpha~ = alpha * keep * drop_scale
    float drop_scale;           // 1 / (1 - p)
    float* score_out;           // [B,n,n] raw Eq.(8) scores s_ij (their sign selects the leaky-relu slope)
    float* alpha_out;           // [B,n,n] softmax weights BEFORE dropout
    uint8_t* relu_mask_out;     // [B,n,D] 1 where (alpha~ h) > 0
    // de-duplicated scoring (all optional): many pairs of one impression share the user graph, so layer 0 of the
    // user graph is projected once per behaviour and every pair reads it through an index
    const int32_t* px_index;    // [B] graph b reads P and X of graph px_index[b] (P then holds K1 WITHOUT k3)
    const int32_t* adj_index;   // [B] graph b reads adj of graph adj_index[b]
    const float* k3; int ldk3;  // [B,ldk3] added to the staged U tile in-kernel: U = fl(K1 + k3), same rounding as the GEMM path
    // node pruning (edge-driven kernel only, optional): row_active [B,n] 0 = node whose output cannot reach any context
    // (digat_user_active_rows): its P row was never computed (may hold anything), no edge of it is evaluated and its
    // output row is NOT written (nothing may read it).  No active node may have an edge to an inactive one.
    const uint8_t* row_active;
    // with row_active: Yc [M_act, D] also receives the output rows of the active nodes, row_pos [B*n] = position of node
    // row r in that compact list -- the next layer's projection GEMM reads Yc directly (no gather pass)
    float* Yc; const int32_t* row_pos;
    // vanilla-GAT scores (ablation encoders, reference graphEncoders.py:498-500 / 515-517; edge-driven kernel only):
    // gat_s [B*n, 2], s_ij = gat_s[j][0] + gat_s[i][1] (= a1 . h_j + a2 . h_i).  P then holds h only (ldp >= D) and `a` is unused.
    const float* gat_s = nullptr;
    // precomputed CSR (digat_build_graph_csr; edge-driven kernel with one graph per CTA): graph b uses record
    // csr_index[b] (or b): rowptr [*, n+1] uint16, meta [*, n*n] uint16.  The adjacency bytes are then not read at all.
    const uint16_t* csr_rowptr = nullptr;
    const uint16_t* csr_meta = nullptr;
    const int32_t* csr_index = nullptr;
};

__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {      // two IEEE fp32 adds in one issue slot
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <bool kSingleTile>
__global__ void __launch_bounds__(kPairThreads, 2)
graph_layer_fwd_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map3,
                       PairAttnArgs p, PairAttnGeom g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];   // TMA destinations need 128-byte alignment
    uint8_t* sbase = smem_raw;                              // (keep it a __shared__ pointer: LDS/STS, not generic LD/ST)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n, D = p.D, R = g.R;
    const int b0 = blockIdx.x * R;
    const int Rv = min(R, p.B - b0);                 // graphs actually present in this CTA
    const int rows = R * n;                          // rows of a staged tile (TMA zero-fills rows past B*n)
    const int src0 = p.px_index != nullptr ? p.px_index[b0] : b0;      // indexed mode runs with R == 1

    const int half = 2 * g.tile_floats;              // floats per pipeline buffer
    float* buf0 = reinterpret_cast<float*>(sbase);   // [2][half]   (128-byte aligned: TMA destination)
    float* a_s = buf0 + 2 * half;                    // [D]
    float* k3_s = a_s + D;                           // [D] (indexed mode only; zero-sized region otherwise is still reserved)
    float* St = k3_s + D;                            // [R][n][lds]   St[r][j*lds + i]
    uint64_t* full = reinterpret_cast<uint64_t*>(St + (size_t)R * n * g.lds);   // [2] TMA barriers (8-byte aligned)

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < D / 4; i += kPairThreads) {
        reinterpret_cast<float4*>(a_s)[i] = reinterpret_cast<const float4*>(p.a)[i];
        if (p.k3 != nullptr)
            reinterpret_cast<float4*>(k3_s)[i] = reinterpret_cast<const float4*>(p.k3 + (size_t)b0 * p.ldk3)[i];
    }
    __syncthreads();

    // unified load schedule: loads 0..nch1-1 are phase-1 chunks (U + K2 tiles), nch1.. are phase-3 chunks (h tile);
    // load l lands in buffer l&1 and completes phase (l>>1)&1 of that buffer's barrier.
    const int n_loads = g.nch1 + g.nch3;
    auto issue = [&](int l) {
        float* dst = buf0 + (l & 1) * half;
        if (l < g.nch1) {
            mbar_arrive_expect_tx(&full[l & 1], 2u * rows * g.dc * 4u);
            tma_load_2d(dst, &map1, &full[l & 1], D + l * g.dc, src0 * n);                     // U  = k3 + K1
            tma_load_2d(dst + g.tile_floats, &map1, &full[l & 1], 2 * D + l * g.dc, src0 * n); // K2
        } else {
            mbar_arrive_expect_tx(&full[l & 1], (uint32_t)rows * g.dc3 * 4u);
            tma_load_2d(dst, &map3, &full[l & 1], (l - g.nch1) * g.dc3, src0 * n);             // h
        }
    };
    if (tid == 0) {
        issue(0);
        if (n_loads > 1) issue(1);
    }

    // ------------------------------------------------------------------ phase 1: pair scores
    const int nt = g.nt, tiles_per_graph = nt * nt, tiles = Rv * tiles_per_graph;
    uint64_t acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = 0ull;

    for (int l = 0; l < g.nch1; ++l) {
        const int c0 = l * g.dc;
        const int wq = min(g.dc, D - c0) >> 2;
        mbar_wait(&full[l & 1], (uint32_t)(l >> 1) & 1u);
        if (p.k3 != nullptr) {                        // indexed mode: the shared K1 tile becomes this pair's U tile
            float* Ut = buf0 + (l & 1) * half;
            for (int it = tid; it < n * wq; it += kPairThreads) {
                const int row = it / wq, q = it - row * wq;
                float4* u = reinterpret_cast<float4*>(Ut + row * g.dc + 4 * q);
                const float4 k = *reinterpret_cast<const float4*>(k3_s + c0 + 4 * q);
                float4 v = *u;
                v.x = k.x + v.x; v.y = k.y + v.y; v.z = k.z + v.z; v.w = k.w + v.w;
                *u = v;
            }
            __syncthreads();
        }
        for (int t = tid; t < tiles; t += kPairThreads) {
            const int r = t / tiles_per_graph, tt = t - r * tiles_per_graph;
            const int ti = tt / nt, tj = tt - ti * nt;
            // byte offsets (from the start of dynamic smem) of this thread's 4 K2 rows and 4 U rows
            const uint32_t buf_off = (uint32_t)((l & 1) * half) * 4u;
            uint32_t ko[4], uo[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                ko[x] = buf_off + (uint32_t)(g.tile_floats + (r * n + min(ti + nt * x, n - 1)) * g.dc) * 4u;
                uo[x] = buf_off + (uint32_t)((r * n + min(tj + nt * x, n - 1)) * g.dc) * 4u;
            }
            if (!kSingleTile) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = 0ull;
            }
            const uint32_t ao = (uint32_t)(2 * half + c0) * 4u;
#pragma unroll 1
            for (int q = 0; q < wq; ++q) {
                const uint32_t qo = (uint32_t)q * 16u;
                const float4 av = *reinterpret_cast<const float4*>(smem_raw + ao + qo);
                const uint64_t a01 = pack2(av.x, av.y), a23 = pack2(av.z, av.w);
                uint64_t u01[4], u23[4];
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const float4 u = *reinterpret_cast<const float4*>(smem_raw + uo[y] + qo);
                    u01[y] = pack2(u.x, u.y);
                    u23[y] = pack2(u.z, u.w);
                }
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const float4 k2 = *reinterpret_cast<const float4*>(smem_raw + ko[x] + qo);
                    const uint64_t k01 = pack2(k2.x, k2.y), k23 = pack2(k2.z, k2.w);
                    // 8 independent add -> relu -> fma chains per x: staged so the scheduler can interleave them
                    float s[16];
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        unpack2(add2(u01[y], k01), s[4 * y + 0], s[4 * y + 1]);
                        unpack2(add2(u23[y], k23), s[4 * y + 2], s[4 * y + 3]);
                    }
#pragma unroll
                    for (int e = 0; e < 16; ++e) s[e] = fmaxf(s[e], 0.f);
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        acc[x][y] = fma2(a01, pack2(s[4 * y + 0], s[4 * y + 1]), acc[x][y]);
                        acc[x][y] = fma2(a23, pack2(s[4 * y + 2], s[4 * y + 3]), acc[x][y]);
                    }
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
                            float e, o;
                            unpack2(acc[x][y], e, o);
                            float* dst = S + j * g.lds + i;
                            *dst = (l == 0 ? 0.f : *dst) + (e + o);
                        }
                    }
            }
        }
        __syncthreads();                              // everyone is done with buffer l&1
        if (tid == 0 && l + 2 < n_loads) issue(l + 2);
    }
    if (kSingleTile && tid < tiles) {
        const int r = tid / tiles_per_graph, tt = tid - r * tiles_per_graph;
        const int ti = tt / nt, tj = tt - ti * nt;
        float* S = St + (size_t)r * n * g.lds;
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const int i = ti + nt * x, j = tj + nt * y;
                if (i < n && j < n) {
                    float e, o;
                    unpack2(acc[x][y], e, o);
                    S[j * g.lds + i] = e + o;
                }
            }
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 2: masked softmax per query node
    for (int row = warp; row < Rv * n; row += kPairThreads / 32) {
        const int r = row / n, i = row - r * n;
        float* S = St + (size_t)r * n * g.lds;
        const size_t ag = p.adj_index != nullptr ? (size_t)p.adj_index[b0 + r] : (size_t)(b0 + r);
        const uint8_t* adj = p.adj + (ag * n + i) * n;
        float v[kPairMaxNodes / 32];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < kPairMaxNodes / 32; ++k) {
            const int j = lane + 32 * k;
            float m = -INFINITY;
            if (j < n) {
                const float s = S[j * g.lds + i];
                if (p.score_out != nullptr) p.score_out[((size_t)(b0 + r) * n + i) * n + j] = s;
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
                float al = v[k] / sum;
                const size_t o = ((size_t)(b0 + r) * n + i) * n + j;
                if (p.alpha_out != nullptr) p.alpha_out[o] = al;
                if (p.drop_keep != nullptr) al = p.drop_keep[o] != 0 ? al * p.drop_scale : 0.f;
                S[j * g.lds + i] = al;
            }
        }
    }
    // padding columns i in [n, lds) are read (and discarded) by phase 3: keep them finite
    {
        const int pad = g.lds - n;
        for (int it = tid; it < Rv * n * pad; it += kPairThreads) {
            const int i = n + it % pad, j = (it / pad) % n, r = it / (pad * n);
            St[(size_t)r * n * g.lds + j * g.lds + i] = 0.f;
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 3: Y = relu(alpha * h) + X
    const int nit = (n + 7) >> 3;
    for (int l = g.nch1; l < n_loads; ++l) {
        const int c0 = (l - g.nch1) * g.dc3;
        const int wq = min(g.dc3, D - c0) >> 2;
        const float* Hs0 = buf0 + (l & 1) * half;
        mbar_wait(&full[l & 1], (uint32_t)(l >> 1) & 1u);
        const int items = Rv * nit * wq;
        for (int it = tid; it < items; it += kPairThreads) {
            const int q = it % wq, ib = (it / wq) % nit, r = it / (wq * nit);
            const int i0 = ib * 8;
            const float* Hs = Hs0 + (size_t)r * n * g.dc3 + 4 * q;
            const float* S = St + (size_t)r * n * g.lds + i0;
            // outer product alpha[8 rows] x h[4 features] with packed fp32x2 FMAs and only 4 register moves per j:
            // row pair (a_k, a_k+1) times the feature pair (h0,h1) gives o[k][0], o[k+1][1]; times the SWAPPED pair
            // (h1,h0) it gives o[k][1], o[k+1][0].  od = "diagonal" accumulators, ox = "crossed" accumulators.
            uint64_t od[4][2], ox[4][2];
#pragma unroll
            for (int k = 0; k < 4; ++k) od[k][0] = od[k][1] = ox[k][0] = ox[k][1] = 0ull;
            {   // the residual rows are needed only after the j loop: pull them into L2 now (no registers held)
                const size_t xrow0 = (p.px_index != nullptr ? (size_t)src0 : (size_t)(b0 + r)) * n + i0;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (i0 + k < n)
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(p.X + (xrow0 + k) * D + c0 + 4 * q));
            }
#pragma unroll 2
            for (int j = 0; j < n; ++j) {
                const float4 h = *reinterpret_cast<const float4*>(Hs);
                const float4 al0 = *reinterpret_cast<const float4*>(S);
                const float4 al1 = *reinterpret_cast<const float4*>(S + 4);
                Hs += g.dc3;
                S += g.lds;
                const uint64_t h01 = pack2(h.x, h.y), h23 = pack2(h.z, h.w);
                const uint64_t h10 = pack2(h.y, h.x), h32 = pack2(h.w, h.z);
                const uint64_t ap[4] = {pack2(al0.x, al0.y), pack2(al0.z, al0.w), pack2(al1.x, al1.y), pack2(al1.z, al1.w)};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    od[k][0] = fma2(ap[k], h01, od[k][0]);      // (a_2k   h0, a_2k+1 h1)
                    ox[k][0] = fma2(ap[k], h10, ox[k][0]);      // (a_2k   h1, a_2k+1 h0)
                    od[k][1] = fma2(ap[k], h23, od[k][1]);      // (a_2k   h2, a_2k+1 h3)
                    ox[k][1] = fma2(ap[k], h32, ox[k][1]);      // (a_2k   h3, a_2k+1 h2)
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float d0, d1, d2, d3, x0, x1, x2, x3;
                unpack2(od[k][0], d0, d1);      // row 2k: h0 ; row 2k+1: h1
                unpack2(ox[k][0], x0, x1);      // row 2k: h1 ; row 2k+1: h0
                unpack2(od[k][1], d2, d3);      // row 2k: h2 ; row 2k+1: h3
                unpack2(ox[k][1], x2, x3);      // row 2k: h3 ; row 2k+1: h2
                const float4 rowv[2] = {make_float4(d0, x0, d2, x2), make_float4(x1, d1, x3, d3)};
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = i0 + 2 * k + e;
                    if (i < n) {
                        const size_t off = ((size_t)(b0 + r) * n + i) * D + c0 + 4 * q;
                        const size_t xoff = p.px_index != nullptr ? ((size_t)src0 * n + i) * D + c0 + 4 * q : off;
                        const float4 x = ldg_stream(reinterpret_cast<const float4*>(p.X + xoff));
                        float4 y = rowv[e];
                        if (p.relu_mask_out != nullptr)
                            *reinterpret_cast<uchar4*>(p.relu_mask_out + off) =
                                make_uchar4(y.x > 0.f, y.y > 0.f, y.z > 0.f, y.w > 0.f);
                        y.x = fmaxf(y.x, 0.f) + x.x;
                        y.y = fmaxf(y.y, 0.f) + x.y;
                        y.z = fmaxf(y.z, 0.f) + x.z;
                        y.w = fmaxf(y.w, 0.f) + x.w;
                        stg_stream(reinterpret_cast<float4*>(p.Y + off), y);
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0 && l + 2 < n_loads) issue(l + 2);
    }
}

// Host-side geometry: graphs per CTA and chunk widths for (n, D) under a shared-memory budget.
inline void pair_attn_geometry(int n, int D, int B, bool indexed, PairAttnGeom* g) {
    const int nt = (n + 3) / 4;
    const int tiles = nt * nt;
    int dc = 68;                                   // 68/4 = 17 (odd): consecutive dense rows start 4 banks apart
    if (dc > D) dc = D;
    const int lds = ((n + 7) / 8) * 8 + 4;         // multiple of 4; +4 shifts consecutive j by 4 banks
    auto tile_floats = [&](int R) { return ((R * n * dc * 4 + 127) / 128) * 128 / 4; };
    auto smem_of = [&](int R) {
        return (size_t)4 * tile_floats(R) * 4 + (size_t)2 * D * 4 + (size_t)R * n * lds * 4 + 16;
    };
    const size_t budget = 110 * 1024;              // two CTAs per SM
    int R = kPairThreads / tiles;
    if (R < 1 || indexed) R = 1;
    while (R > 1 && (smem_of(R) > budget || R * n > 256)) --R;      // TMA box: at most 256 rows
    if (R > B) R = B > 0 ? B : 1;
    g->R = R; g->nt = nt; g->dc = dc; g->nch1 = (D + dc - 1) / dc; g->dc3 = 2 * dc;
    g->nch3 = (D + 2 * dc - 1) / (2 * dc); g->lds = lds; g->tile_floats = tile_floats(R); g->smem = smem_of(R);
}

struct PairAttnArgs;
int launch_graph_layer_fwd_sparse(const PairAttnArgs& args, int n_src, cudaStream_t st);    // pair_attention_sparse.cuh
size_t graph_layer_fwd_sparse_smem(int n, int D);                                            // pair_attention_sparse.cuh
static int g_layer_mode = 0;   // 0 = auto (edge-driven kernel for single-graph CTAs at inference), 1 = dense, 2 = edge-driven

inline int launch_graph_layer_fwd(const float* P, int ldp, const float* a, const uint8_t* adj, const float* X, float* Y,
                                  int B, int n, int D, const uint8_t* drop_keep, float drop_scale, float* score_out,
                                  float* alpha_out, uint8_t* relu_mask_out, const int32_t* px_index, int n_src,
                                  const int32_t* adj_index, const float* k3, int ldk3, const uint8_t* row_active,
                                  float* Yc, const int32_t* row_pos, const uint16_t* csr_rowptr, const uint16_t* csr_meta,
                                  const int32_t* csr_index, cudaStream_t st) {
    if (B == 0) return DIGAT_OK;
    DIGAT_REQUIRE((csr_rowptr == nullptr) == (csr_meta == nullptr), "digat_graph_layer_fwd: csr_rowptr and csr_meta go together");
    DIGAT_REQUIRE((Yc == nullptr) == (row_pos == nullptr) && (Yc == nullptr || (row_active != nullptr && aligned16(Yc))),
                  "digat_graph_layer_fwd: Yc, row_pos and row_active go together");
    DIGAT_REQUIRE(P && a && adj && X && Y, "digat_graph_layer_fwd: null pointer");
    DIGAT_REQUIRE(B >= 0 && n >= 1 && n <= kPairMaxNodes, "digat_graph_layer_fwd: n=%d outside [1,%d]", n, kPairMaxNodes);
    DIGAT_REQUIRE(D >= 4 && (D & 3) == 0 && D <= 1024, "digat_graph_layer_fwd: D=%d must be a multiple of 4 in [4,1024]", D);
    DIGAT_REQUIRE((ldp & 3) == 0 && ldp >= 3 * D, "digat_graph_layer_fwd: ldp=%d must be a multiple of 4 and >= 3D", ldp);
    DIGAT_REQUIRE(aligned16(P) && aligned16(a) && aligned16(X) && aligned16(Y),
                  "digat_graph_layer_fwd: pointers must be 16-byte aligned");
    DIGAT_REQUIRE((px_index == nullptr) == (k3 == nullptr) && (px_index == nullptr || n_src > 0),
                  "digat_graph_layer_fwd: px_index, k3 and n_src go together");
    DIGAT_REQUIRE(k3 == nullptr || (aligned16(k3) && (ldk3 & 3) == 0 && ldk3 >= D),
                  "digat_graph_layer_fwd: k3 must be 16-byte aligned with ldk3 a multiple of 4 and >= D");
    PairAttnGeom g;
    pair_attn_geometry(n, D, B, px_index != nullptr, &g);
    const int64_t src_graphs = px_index != nullptr ? n_src : B;
    const DeviceInfo* di = device_info();
    if (!di) return fail(DIGAT_E_CUDA, "digat_graph_layer_fwd: no CUDA device");
    DIGAT_REQUIRE(g.smem <= (size_t)di->max_smem_optin, "digat_graph_layer_fwd: needs %zu B shared memory", g.smem);
    CUtensorMap map1, map3;
    int rc;
    if ((rc = make_tensor_map_2d(&map1, P, src_graphs * n, 3 * D, ldp, g.R * n, g.dc, CU_TENSOR_MAP_SWIZZLE_NONE)) != DIGAT_OK) return rc;
    if ((rc = make_tensor_map_2d(&map3, P, src_graphs * n, 3 * D, ldp, g.R * n, g.dc3, CU_TENSOR_MAP_SWIZZLE_NONE)) != DIGAT_OK) return rc;
    PairAttnArgs args{P, ldp, a, adj, X, Y, B, n, D, drop_keep, drop_scale, score_out, alpha_out, relu_mask_out,
                      px_index, adj_index, k3, ldk3, row_active, Yc, row_pos};
    args.csr_rowptr = csr_rowptr;
    args.csr_meta = csr_meta;
    args.csr_index = csr_index;
    const bool inference = !drop_keep && !score_out && !alpha_out && !relu_mask_out;
    // Inference takes the edge-driven kernel (small graphs are batched several per CTA); the dense kernel stays for graphs
    // whose edge-driven working set does not fit one CTA (n > ~100 at D = 400) and for training WITHOUT a precomputed CSR.
    // Training with csr_rowptr / csr_meta also takes the edge-driven kernel: score_out / alpha_out are then PER-EDGE arrays in
    // CSR order ([B, n*n] capacity each) for digat_graph_layer_bwd_csr, not dense [B,n,n] matrices.
    const bool sparse_fits = graph_layer_fwd_sparse_smem(n, D) <= (size_t)di->max_smem_optin;
    const bool sparse_train = !inference && csr_rowptr != nullptr && px_index == nullptr && g_layer_mode != 1 && sparse_fits;
    if ((inference && g_layer_mode != 1 && (sparse_fits || g_layer_mode == 2)) || sparse_train)
        return launch_graph_layer_fwd_sparse(args, n_src, st);
    if (!inference && csr_rowptr != nullptr)
        return fail(DIGAT_E_UNSUPPORTED, "digat_graph_layer_fwd: training with a CSR needs the edge-driven kernel (not indexed, "
                    "working set within one SM); query digat_graph_layer_csr_training_supported first");
    if (row_active != nullptr)
        return fail(DIGAT_E_UNSUPPORTED, "digat_graph_layer_fwd: row_active needs the edge-driven kernel (inference, one graph "
                    "per CTA); query digat_graph_layer_supports_row_active first");
    const int grid = (B + g.R - 1) / g.R;
    const bool single = g.R * g.nt * g.nt <= kPairThreads;
    if (single) {
        if (int rc_ = ensure_dynamic_smem(graph_layer_fwd_kernel<true>, (size_t)(g.smem))) return rc_;
        graph_layer_fwd_kernel<true><<<grid, kPairThreads, g.smem, st>>>(map1, map3, args, g);
    } else {
        if (int rc_ = ensure_dynamic_smem(graph_layer_fwd_kernel<false>, (size_t)(g.smem))) return rc_;
        graph_layer_fwd_kernel<false><<<grid, kPairThreads, g.smem, st>>>(map1, map3, args, g);
    }
    return check_launch("digat_graph_layer_fwd");
}

// 1 when launch_graph_layer_fwd would take the edge-driven kernel for an inference call with these sizes (the only
// kernel that honours row_active): one graph per CTA (not indexed) and its working set fits the SM.
inline int graph_layer_supports_row_active(int n, int D, int B) {
    if (n < 1 || n > kPairMaxNodes || D < 4 || (D & 3) != 0 || D > 1024 || B < 1 || g_layer_mode == 1) return 0;
    const DeviceInfo* di = device_info();
    if (!di) return 0;
    return graph_layer_fwd_sparse_smem(n, D) <= (size_t)di->max_smem_optin ? 1 : 0;
}

}  // namespace digat
