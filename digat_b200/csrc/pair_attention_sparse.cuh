// Fused Eq. (8) graph-attention layer, forward, EDGE-DRIVEN variant (inference; one graph per CTA).
//
// The reference evaluates s_ij for all n^2 pairs and then overwrites the non-edges with -1e9
// (graphEncoders.py:150-152).  A masked pair contributes exp(-1e9 - max) == 0.0f to the softmax of any row that has at
// least one edge, so its score is never observable: this kernel computes s_ij ONLY on the edges of the adjacency and
// aggregates over neighbours only -- bit-for-bit the same alpha and Y as the dense evaluation, with E instead of n^2
// pair evaluations (MIND-shaped user graphs: E ~ 300-650 of 4624; SAG trees: ~2n of n^2).  A row without any edge
// (cannot happen in the reference's data: the diagonal is always set) is handled like the reference: every entry is
// -1e9, the softmax is uniform 1/n over ALL nodes.
//
// Structure (same 2-deep TMA pipeline and smem tiles as the dense kernel, pair_attention.cuh):
//   setup    adjacency bytes -> CSR (row pointers, uint8 columns) with warp ballots + one warp scan
//   phase 1  work item = (row i, up to 4 neighbours): K2_i quad in registers, U_j quads from smem, packed
//            FADD2 / FMNMX / FFMA2; partial dot products accumulate into score[e] in smem once per feature chunk
//   phase 2  one warp per row: leaky-relu, max/sum by shuffles over the row's CSR segment, exp, normalise
//   phase 3  work item = (row i, feature quad): loop over the neighbours, FFMA2; relu + residual, streaming store
// HBM traffic per graph is unchanged (P is still streamed once); the kernel moves from fp32-pipe-bound to HBM-bound.
#pragma once
#include "common.cuh"
#include "tma.cuh"
#include "pair_attention.cuh"

namespace digat {

constexpr int kSparseItemEdges = 4;

struct SparseGeom {
    int dc, nch1, dc3, nch3;
    int tile_floats;       // floats of one [n][dc] tile rounded up to 128 bytes
    int max_items;         // upper bound of phase-1 work items: sum_i ceil(deg_i / 4) <= n * ceil(n/4)
    size_t smem;
};

template <int kDummy>
__global__ void __launch_bounds__(kPairThreads, 2)
graph_layer_fwd_sparse_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map3,
                              PairAttnArgs p, SparseGeom g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n, D = p.D;
    const int b = blockIdx.x;
    const int src0 = p.px_index != nullptr ? p.px_index[b] : b;
    const int half = 2 * g.tile_floats;

    float* buf0 = reinterpret_cast<float*>(smem_raw);             // [2][half]   (multiple of 128 bytes)
    uint64_t* full = reinterpret_cast<uint64_t*>(buf0 + 2 * half); // [2] TMA barriers
    float* a_s = reinterpret_cast<float*>(full + 2);               // [D]
    float* k3_s = a_s + D;                                         // [D]
    float* score = k3_s + D;                                       // [n*n] per-edge score, later alpha~
    int* rowptr = reinterpret_cast<int*>(score + n * n);           // [n+1]
    int* itemptr = rowptr + (n + 1);                               // [n+1] first phase-1 item of each row
    int* items = itemptr + (n + 1);                                // [max_items] (row << 16) | first edge offset in row
    uint8_t* col = reinterpret_cast<uint8_t*>(items + g.max_items); // [n*n] neighbour index of each edge
    uint8_t* adj_s = col + n * n;                                  // [n*n] adjacency bytes
    uint8_t* uniform_row = adj_s + n * n;                          // [n] 1 = row without edges (uniform softmax)

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < D / 4; i += kPairThreads) {
        reinterpret_cast<float4*>(a_s)[i] = reinterpret_cast<const float4*>(p.a)[i];
        if (p.k3 != nullptr)
            reinterpret_cast<float4*>(k3_s)[i] = reinterpret_cast<const float4*>(p.k3 + (size_t)b * p.ldk3)[i];
    }
    {
        const size_t ag = p.adj_index != nullptr ? (size_t)p.adj_index[b] : (size_t)b;
        const uint8_t* srcp = p.adj + ag * n * n;
        for (int i = tid; i < n * n; i += kPairThreads) adj_s[i] = srcp[i];
    }
    __syncthreads();

    const int n_loads = g.nch1 + g.nch3;
    auto issue = [&](int l) {
        float* dst = buf0 + (l & 1) * half;
        if (l < g.nch1) {
            mbar_arrive_expect_tx(&full[l & 1], 2u * n * g.dc * 4u);
            tma_load_2d(dst, &map1, &full[l & 1], D + l * g.dc, src0 * n);
            tma_load_2d(dst + g.tile_floats, &map1, &full[l & 1], 2 * D + l * g.dc, src0 * n);
        } else {
            mbar_arrive_expect_tx(&full[l & 1], (uint32_t)n * g.dc3 * 4u);
            tma_load_2d(dst, &map3, &full[l & 1], (l - g.nch1) * g.dc3, src0 * n);
        }
    };
    if (tid == 0) {
        issue(0);
        if (n_loads > 1) issue(1);
    }

    // ------------------------------------------------------------------ CSR of the adjacency (overlaps the first TMA loads)
    // pass A: degrees (a row without edges becomes a full row with uniform weights)
    for (int i = warp; i < n; i += kPairThreads / 32) {
        int deg = 0;
        for (int k = 0; k < (n + 31) / 32; ++k) {
            const int j = lane + 32 * k;
            deg += __popc(__ballot_sync(0xffffffffu, j < n && adj_s[i * n + j] != 0));
        }
        if (lane == 0) {
            uniform_row[i] = deg == 0;
            if (deg == 0) deg = n;
            rowptr[i + 1] = deg;                                   // degrees for now, scanned below
            itemptr[i + 1] = (deg + kSparseItemEdges - 1) / kSparseItemEdges;
        }
    }
    __syncthreads();
    if (warp == 0) {                                               // inclusive scan of <= 128 entries by one warp
        int run_e = 0, run_i = 0;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            int ve = i < n ? rowptr[i + 1] : 0, vi = i < n ? itemptr[i + 1] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int te = __shfl_up_sync(0xffffffffu, ve, o), ti = __shfl_up_sync(0xffffffffu, vi, o);
                if (lane >= o) { ve += te; vi += ti; }
            }
            if (i < n) { rowptr[i + 1] = run_e + ve; itemptr[i + 1] = run_i + vi; }
            run_e += __shfl_sync(0xffffffffu, ve, 31);
            run_i += __shfl_sync(0xffffffffu, vi, 31);
        }
        if (lane == 0) { rowptr[0] = 0; itemptr[0] = 0; }
    }
    __syncthreads();
    // pass B: columns, phase-1 items, zeroed scores
    for (int i = warp; i < n; i += kPairThreads / 32) {
        const int e0 = rowptr[i];
        const bool uni = uniform_row[i] != 0;
        int filled = 0;
        for (int k = 0; k < (n + 31) / 32; ++k) {
            const int j = lane + 32 * k;
            const bool on = j < n && (uni || adj_s[i * n + j] != 0);
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (on) col[e0 + filled + __popc(m & ((1u << lane) - 1u))] = (uint8_t)j;
            filled += __popc(m);
        }
        const int it0 = itemptr[i], nit = itemptr[i + 1] - it0;
        for (int t = lane; t < nit; t += 32) items[it0 + t] = (i << 16) | (t * kSparseItemEdges);
    }
    const int E = rowptr[n], n_items = itemptr[n];
    for (int e = tid; e < E; e += kPairThreads) score[e] = 0.f;
    __syncthreads();

    // ------------------------------------------------------------------ phase 1: scores on the edges only
    for (int l = 0; l < g.nch1; ++l) {
        const int c0 = l * g.dc;
        const int wq = min(g.dc, D - c0) >> 2;
        float* Ut = buf0 + (l & 1) * half;
        mbar_wait(&full[l & 1], (uint32_t)(l >> 1) & 1u);
        if (p.k3 != nullptr) {                                     // indexed mode: shared K1 tile -> this pair's U tile
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
        const uint32_t uoff = (uint32_t)((l & 1) * half) * 4u;
        const uint32_t koff = uoff + (uint32_t)g.tile_floats * 4u;
        const uint32_t aoff = (uint32_t)(2 * half + 4 + c0) * 4u;     // a_s sits after the two 8-byte barriers
        for (int it = tid; it < n_items; it += kPairThreads) {
            const int packed = items[it];
            const int i = packed >> 16, eo = packed & 0xffff;
            const int e0 = rowptr[i] + eo;
            const int cnt = min(kSparseItemEdges, rowptr[i + 1] - e0);
            uint32_t uo[kSparseItemEdges];
#pragma unroll
            for (int m = 0; m < kSparseItemEdges; ++m)
                uo[m] = uoff + (uint32_t)(col[e0 + min(m, cnt - 1)] * g.dc) * 4u;      // clamp: idle slots redo the last edge
            const uint32_t ko = koff + (uint32_t)(i * g.dc) * 4u;
            uint64_t acc[kSparseItemEdges];
#pragma unroll
            for (int m = 0; m < kSparseItemEdges; ++m) acc[m] = 0ull;
#pragma unroll 1
            for (int q = 0; q < wq; ++q) {
                const uint32_t qo = (uint32_t)q * 16u;
                const float4 av = *reinterpret_cast<const float4*>(smem_raw + aoff + qo);
                const float4 k2 = *reinterpret_cast<const float4*>(smem_raw + ko + qo);
                const uint64_t a01 = pack2(av.x, av.y), a23 = pack2(av.z, av.w);
                const uint64_t k01 = pack2(k2.x, k2.y), k23 = pack2(k2.z, k2.w);
                float s[4 * kSparseItemEdges];
#pragma unroll
                for (int m = 0; m < kSparseItemEdges; ++m) {
                    const float4 u = *reinterpret_cast<const float4*>(smem_raw + uo[m] + qo);
                    unpack2(add2(pack2(u.x, u.y), k01), s[4 * m + 0], s[4 * m + 1]);
                    unpack2(add2(pack2(u.z, u.w), k23), s[4 * m + 2], s[4 * m + 3]);
                }
#pragma unroll
                for (int e = 0; e < 4 * kSparseItemEdges; ++e) s[e] = fmaxf(s[e], 0.f);
#pragma unroll
                for (int m = 0; m < kSparseItemEdges; ++m) acc[m] = fma2(a01, pack2(s[4 * m + 0], s[4 * m + 1]), acc[m]);
#pragma unroll
                for (int m = 0; m < kSparseItemEdges; ++m) acc[m] = fma2(a23, pack2(s[4 * m + 2], s[4 * m + 3]), acc[m]);
            }
#pragma unroll
            for (int m = 0; m < kSparseItemEdges; ++m)
                if (m < cnt) {
                    float lo, hi;
                    unpack2(acc[m], lo, hi);
                    score[e0 + m] += lo + hi;                      // each edge is owned by exactly one thread
                }
        }
        __syncthreads();
        if (tid == 0 && l + 2 < n_loads) issue(l + 2);
    }

    // ------------------------------------------------------------------ phase 2: softmax over each row's edges
    for (int i = warp; i < n; i += kPairThreads / 32) {
        const int e0 = rowptr[i], deg = rowptr[i + 1] - e0;
        const bool uni = uniform_row[i] != 0;
        float v[kPairMaxNodes / 32];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < kPairMaxNodes / 32; ++k) {
            const int t = lane + 32 * k;
            float m = -INFINITY;
            if (t < deg) {
                const float s = score[e0 + t];
                m = uni ? kNegFill : (s > 0.f ? s : s * kLeakySlope);
            }
            v[k] = m;
            mx = fmaxf(mx, m);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < kPairMaxNodes / 32; ++k) {
            const int t = lane + 32 * k;
            v[k] = (t < deg) ? expf(v[k] - mx) : 0.f;
            sum += v[k];
        }
        sum = warp_sum(sum);
#pragma unroll
        for (int k = 0; k < kPairMaxNodes / 32; ++k) {
            const int t = lane + 32 * k;
            if (t < deg) score[e0 + t] = v[k] / sum;
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 3: Y = relu(sum_{j in N(i)} alpha_ij h_j) + X
    for (int l = g.nch1; l < n_loads; ++l) {
        const int c0 = (l - g.nch1) * g.dc3;
        const int wq = min(g.dc3, D - c0) >> 2;
        const float* Hs0 = buf0 + (l & 1) * half;
        mbar_wait(&full[l & 1], (uint32_t)(l >> 1) & 1u);
        for (int it = tid; it < n * wq; it += kPairThreads) {
            const int i = it / wq, q = it - i * wq;
            const size_t yoff = ((size_t)b * n + i) * D + c0 + 4 * q;
            const size_t xoff = ((size_t)src0 * n + i) * D + c0 + 4 * q;
            const float4 x = ldg_stream(reinterpret_cast<const float4*>(p.X + xoff));      // consumed after the loop
            const float* Hs = Hs0 + 4 * q;
            uint64_t o01 = 0ull, o23 = 0ull;
            const int e1 = rowptr[i + 1];
            for (int e = rowptr[i]; e < e1; ++e) {
                const float al = score[e];
                const float4 h = *reinterpret_cast<const float4*>(Hs + col[e] * g.dc3);
                const uint64_t aa = pack2(al, al);
                o01 = fma2(aa, pack2(h.x, h.y), o01);
                o23 = fma2(aa, pack2(h.z, h.w), o23);
            }
            float4 y;
            unpack2(o01, y.x, y.y);
            unpack2(o23, y.z, y.w);
            y.x = fmaxf(y.x, 0.f) + x.x;
            y.y = fmaxf(y.y, 0.f) + x.y;
            y.z = fmaxf(y.z, 0.f) + x.z;
            y.w = fmaxf(y.w, 0.f) + x.w;
            stg_stream(reinterpret_cast<float4*>(p.Y + yoff), y);
        }
        __syncthreads();
        if (tid == 0 && l + 2 < n_loads) issue(l + 2);
    }
}

inline void sparse_geometry(int n, int D, SparseGeom* g) {
    int dc = 68;
    if (dc > D) dc = D;
    g->dc = dc; g->nch1 = (D + dc - 1) / dc; g->dc3 = 2 * dc; g->nch3 = (D + 2 * dc - 1) / (2 * dc);
    g->tile_floats = ((n * dc * 4 + 127) / 128) * 128 / 4;
    g->max_items = n * ((n + kSparseItemEdges - 1) / kSparseItemEdges);
    g->smem = (size_t)4 * g->tile_floats * 4 + 16 + (size_t)2 * D * 4 + (size_t)n * n * 4 + (size_t)2 * (n + 1) * 4 +
              (size_t)g->max_items * 4 + (size_t)2 * n * n + (size_t)n + 16;
}

int launch_graph_layer_fwd_sparse(const PairAttnArgs& args, int n_src, cudaStream_t st) {
    SparseGeom g;
    sparse_geometry(args.n, args.D, &g);
    const DeviceInfo* di = device_info();
    if (!di) return fail(DIGAT_E_CUDA, "digat_graph_layer_fwd: no CUDA device");
    DIGAT_REQUIRE(g.smem <= (size_t)di->max_smem_optin, "digat_graph_layer_fwd(sparse): needs %zu B shared memory", g.smem);
    CUtensorMap map1, map3;
    int rc;
    const int64_t src_graphs = args.px_index != nullptr ? n_src : args.B;
    if ((rc = make_tensor_map_2d(&map1, args.P, src_graphs * args.n, 3 * args.D, args.ldp, args.n, g.dc, CU_TENSOR_MAP_SWIZZLE_NONE)) != DIGAT_OK) return rc;
    if ((rc = make_tensor_map_2d(&map3, args.P, src_graphs * args.n, 3 * args.D, args.ldp, args.n, g.dc3, CU_TENSOR_MAP_SWIZZLE_NONE)) != DIGAT_OK) return rc;
    DIGAT_CUDA(cudaFuncSetAttribute(graph_layer_fwd_sparse_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    graph_layer_fwd_sparse_kernel<0><<<args.B, kPairThreads, g.smem, st>>>(map1, map3, args, g);
    return check_launch("digat_graph_layer_fwd(sparse)");
}

}  // namespace digat
