// Backward of the fused Eq. (8) graph-attention layer, EDGE-DRIVEN (training with a precomputed CSR + its transpose).
//
// The dense backward (pair_attention_bwd.cuh) evaluates all n^2 pairs of a graph; like the forward, only the E edges matter:
// a masked pair has alpha = 0 and ds = 0.  With the forward's per-edge outputs (raw score s_e, softmax weight alpha_e, CSR
// order) and G = dY * 1[Z > 0]:
//   pass A  (G, h tiles)   dalpha~_e = G_i . h_j            one edge per thread, accumulated over the feature chunks
//                          dh_j      = sum_{e=(i,j)} alpha~_e G_i       8 lanes per node walk the node's COLUMN (CSC)
//   phase B (per row)      dalpha = dalpha~ * keep * scale;  dm_e = alpha_e (dalpha_e - sum_k alpha_k dalpha_k);
//                          ds_e = dm_e * (s_e > 0 ? 1 : 0.2)            (0 on an edge-less "uniform" row: every entry masked)
//   pass C  (U, K2 tiles)  x_ed = U_jd + K2_id  (the forward's IEEE add on the same operands: the same relu mask)
//                          dK2_id = a_d sum_{e in row i}    ds_e 1[x_ed > 0]      8 lanes per node, its ROW
//                          dU_jd  = a_d sum_{e in column j} ds_e 1[x_ed > 0]      8 lanes per node, its COLUMN
//                          da_d   = sum_e ds_e relu(x_ed)                          per-warp partials [B*10, D]
// Optional extras for the fused projection+layer backward: the saved relu mask of Z is applied to the staged dY tile (no
// separate G = dY * mask pass), and the column sums of dh and dU come out as per-warp partials [B*10, D] (the bias gradient
// is their sum over all rows; the sum over a graph's 10 rows of the dU partials IS dk3) so that nobody has to re-read the
// [B*n, 3D] dP for them.
// Same skeleton as the forward: one CTA per graph, a producer warp streams [n][32]-feature tile pairs through a TMA ring
// (SWIZZLE_128B), consumer warps synchronise through the ring's mbarriers.  Every sum runs in a fixed order (deterministic).
#pragma once
#include "common.cuh"
#include "tma.cuh"
#include "pair_attention.cuh"
#include "pair_attention_sparse.cuh"

namespace digat {

constexpr int kSbwdMaxBufs = 3;
constexpr int kSbwdDc = 32;
constexpr int kSbwdParts = 10;        // rows per graph of da_partial / dh_sum / du_sum: one per consumer warp

struct SparseBwdArgs {
    const float* P; int ldp; const float* a; const float* G;
    const uint8_t* relu_mask;      // [B,n,D] or null: G is dY and the saved mask of Z > 0 is applied to the staged tile
    float* dh_sum; float* du_sum;  // [B*kSbwdParts, D] each or null: per-warp column sums of dh (-> bias gradient) and dU (-> dk3)
    float* gat_ds12;               // [B*n, 2] or null.  Non-null = vanilla-GAT layer (e_ij = leaky_relu(s1_j + s2_i)): P holds h only,
                                   // pass C is replaced by ds1_j = sum over column j of ds_e, ds2_i = sum over row i of ds_e
    const uint16_t* rowptr; const uint16_t* meta; const uint16_t* colptr; const uint16_t* cedge;
    const float* e_score; const float* e_alpha; const uint8_t* drop_keep; float drop_scale;
    float* dP; int lddp; float* da_partial;
    int B, n, D;
};

struct SparseBwdGeom {
    int nch;             // feature chunks of 32
    int tile_floats;     // floats of one [n][32] tile rounded up to 1024 bytes
    int bufs;            // ring depth: 3, or 2 when that lets two CTAs share an SM
    size_t smem;
};

__global__ void __launch_bounds__(kSparseThreads, 2)
graph_layer_bwd_sparse_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapP,
                              SparseBwdArgs p, SparseBwdGeom g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n, D = p.D, b = blockIdx.x;
    const int unit_floats = 2 * g.tile_floats;
    const int kSbwdBufs = g.bufs;
    float* ring = reinterpret_cast<float*>(smem_raw);                       // [bufs][2 tiles]
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + kSbwdBufs * unit_floats);
    uint64_t* empty = full + kSbwdMaxBufs;
    float* a_s = reinterpret_cast<float*>(empty + kSbwdMaxBufs);               // [D]
    float* alt = a_s + D;                                                   // [n*n] alpha~ per edge
    float* dal = alt + n * n;                                               // [n*n] dalpha~ per edge, later ds
    int* rowptr = reinterpret_cast<int*>(dal + n * n);                      // [n+1]
    int* colptr = rowptr + (n + 1);                                         // [n+1]
    uint16_t* meta = reinterpret_cast<uint16_t*>(colptr + (n + 1));         // [n*n]
    uint16_t* cedge = meta + n * n;                                         // [n*n]
    uint8_t* uniform_row = reinterpret_cast<uint8_t*>(cedge + n * n);       // [n]

    if (tid == 0) {
        for (int i = 0; i < kSbwdBufs; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kSparseConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool gat = p.gat_ds12 != nullptr;
    const int n_loads = gat ? g.nch : 2 * g.nch;

    if (warp == kSparseConsumers / 32) {
        // ------------------------------------------------------------------ producer: (G, h) chunks, then (U, K2) chunks
        if (lane == 0) {
            for (int l = 0; l < n_loads; ++l) {
                const int buf = l % kSbwdBufs;
                float* dst = ring + buf * unit_floats;
                mbar_wait(&empty[buf], ((uint32_t)(l / kSbwdBufs) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&full[buf], 2u * n * kSbwdDc * 4u);
                if (l < g.nch) {
                    tma_load_2d(dst, &mapG, &full[buf], l * kSbwdDc, b * n);                               // G
                    tma_load_2d(dst + g.tile_floats, &mapP, &full[buf], l * kSbwdDc, b * n);               // h
                } else {
                    tma_load_2d(dst, &mapP, &full[buf], D + (l - g.nch) * kSbwdDc, b * n);                 // U
                    tma_load_2d(dst + g.tile_floats, &mapP, &full[buf], 2 * D + (l - g.nch) * kSbwdDc, b * n);   // K2
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- setup: CSR + transpose, alpha~, zeroed accumulators
    if (!gat)
        for (int i = tid; i < D; i += kSparseConsumers) a_s[i] = p.a[i];
    {
        const uint16_t* rp = p.rowptr + (size_t)b * (n + 1);
        const uint16_t* cp = p.colptr + (size_t)b * (n + 1);
        for (int i = tid; i <= n; i += kSparseConsumers) {
            const uint32_t raw = rp[i];
            rowptr[i] = (int)(raw & 0x7fffu);
            if (i > 0) uniform_row[i - 1] = (uint8_t)(raw >> 15);
            colptr[i] = (int)cp[i];
        }
    }
    consumer_sync();
    const int E = rowptr[n];
    {
        const uint16_t* mg = p.meta + (size_t)b * n * n;
        const uint16_t* cg = p.cedge + (size_t)b * n * n;
        const float* ea = p.e_alpha + (size_t)b * n * n;
        const uint8_t* keep = p.drop_keep != nullptr ? p.drop_keep + (size_t)b * n * n : nullptr;
        for (int e = tid; e < E; e += kSparseConsumers) {
            const uint32_t mt = mg[e];
            meta[e] = (uint16_t)mt;
            cedge[e] = cg[e];
            float al = ea[e];
            if (keep != nullptr) al = keep[(mt >> 8) * n + (mt & 255u)] != 0 ? al * p.drop_scale : 0.f;
            alt[e] = al;
            dal[e] = 0.f;
        }
    }
    consumer_sync();

    const int q8 = lane & 7, grp = lane >> 3;                       // 8 lanes (feature quads of a 32-wide chunk) per node, 4 nodes per warp
    constexpr int kNodeStep = (kSparseConsumers / 32) * 4;          // nodes per pass of the 10 consumer warps
    auto swz = [](uint32_t base, uint32_t row, uint32_t q) {        // byte address of quad q of row `row` in a SWIZZLE_128B tile
        const uint32_t ro = base + row * (uint32_t)(kSbwdDc * 4);
        return ro + ((q * 16u) ^ (((ro >> 7) & 7u) << 4));
    };

    // Per-graph column sums (dh -> bias gradient, dU -> dk3, da) without atomics and without a barrier: a thread sums its
    // nodes, the four node groups of a warp are folded with shuffles and lanes 0..7 write the WARP's partial to row
    // b * kSbwdParts + warp of the output; the caller adds the kSbwdParts rows of a graph (digat_groupsum) or all rows
    // (digat_colsum).  (Folding the warps here cost a barrier over all consumer warps per unit: 30 % of the stall samples.)
    constexpr int kWarps = kSparseConsumers / 32;
    static_assert(kWarps == kSbwdParts, "partial rows per graph");
    auto park = [&](float* dst, int c0, int wq, float4 v) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, 8);  v.y += __shfl_xor_sync(0xffffffffu, v.y, 8);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, 8);  v.w += __shfl_xor_sync(0xffffffffu, v.w, 8);
        v.x += __shfl_xor_sync(0xffffffffu, v.x, 16); v.y += __shfl_xor_sync(0xffffffffu, v.y, 16);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, 16); v.w += __shfl_xor_sync(0xffffffffu, v.w, 16);
        if (lane < wq) *reinterpret_cast<float4*>(dst + ((size_t)b * kWarps + warp) * D + c0 + 4 * lane) = v;   // wq <= 8
    };

    // ---------------------------------------------------------------------- pass A: dalpha~ per edge, dh per node
    for (int l = 0; l < g.nch; ++l) {
        const int buf = l % kSbwdBufs;
        const int c0 = l * kSbwdDc;
        const int wq = min(kSbwdDc, D - c0) >> 2;
        const uint32_t goff = (uint32_t)(buf * unit_floats) * 4u, hoff = goff + (uint32_t)g.tile_floats * 4u;
        mbar_wait(&full[buf], (uint32_t)(l / kSbwdBufs) & 1u);
        if (p.relu_mask != nullptr) {                               // G = dY * 1[Z > 0], applied to the staged dY tile
            const uint8_t* mk = p.relu_mask + (size_t)b * n * D + c0;
            for (int idx = tid; idx < n * 8; idx += kSparseConsumers) {
                const int row = idx >> 3, q = idx & 7;
                if (q < wq) {
                    const uchar4 m = *reinterpret_cast<const uchar4*>(mk + (size_t)row * D + 4 * q);
                    float4* cell = reinterpret_cast<float4*>(smem_raw + swz(goff, row, q));
                    float4 v = *cell;
                    v.x = m.x ? v.x : 0.f; v.y = m.y ? v.y : 0.f; v.z = m.z ? v.z : 0.f; v.w = m.w ? v.w : 0.f;
                    *cell = v;
                }
            }
            consumer_sync();
        }
        for (int e = tid; e < E; e += kSparseConsumers) {
            const uint32_t mt = meta[e];
            float acc = 0.f;
#pragma unroll 4
            for (int q = 0; q < wq; ++q) {
                const float4 gi = *reinterpret_cast<const float4*>(smem_raw + swz(goff, mt >> 8, q));
                const float4 hj = *reinterpret_cast<const float4*>(smem_raw + swz(hoff, mt & 255u, q));
                acc = fmaf(gi.x, hj.x, acc); acc = fmaf(gi.y, hj.y, acc); acc = fmaf(gi.z, hj.z, acc); acc = fmaf(gi.w, hj.w, acc);
            }
            dal[e] += acc;
        }
        float4 dh_tot = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = warp * 4 + grp; j < n; j += kNodeStep) {       // dh_j = sum over the incoming edges (column j)
            if (q8 < wq) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int ce = colptr[j]; ce < colptr[j + 1]; ++ce) {
                    const int e = cedge[ce];
                    const float al = alt[e];
                    const float4 gi = *reinterpret_cast<const float4*>(smem_raw + swz(goff, meta[e] >> 8, q8));
                    acc.x = fmaf(al, gi.x, acc.x); acc.y = fmaf(al, gi.y, acc.y); acc.z = fmaf(al, gi.z, acc.z); acc.w = fmaf(al, gi.w, acc.w);
                }
                *reinterpret_cast<float4*>(p.dP + ((size_t)b * n + j) * p.lddp + c0 + 4 * q8) = acc;
                dh_tot.x += acc.x; dh_tot.y += acc.y; dh_tot.z += acc.z; dh_tot.w += acc.w;
            }
        }
        if (p.relu_mask != nullptr) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // our writes, then TMA's
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[buf]);
        if (p.dh_sum != nullptr) park(p.dh_sum, c0, wq, dh_tot);
    }
    consumer_sync();

    // ---------------------------------------------------------------------- phase B: softmax / leaky-relu backward per row
    {
        const float* ea = p.e_alpha + (size_t)b * n * n;
        const float* es = p.e_score + (size_t)b * n * n;
        const uint8_t* keep = p.drop_keep != nullptr ? p.drop_keep + (size_t)b * n * n : nullptr;
        for (int i = warp; i < n; i += kSparseConsumers / 32) {
            const int e0 = rowptr[i], deg = rowptr[i + 1] - e0;
            if (deg == 0) continue;
            const bool uni = uniform_row[i] != 0;
            float t = 0.f;
            for (int k = lane; k < deg; k += 32) {
                const int e = e0 + k;
                float da = dal[e];
                if (keep != nullptr) da = keep[i * n + (meta[e] & 255u)] != 0 ? da * p.drop_scale : 0.f;
                dal[e] = da;                                        // dalpha
                t = fmaf(ea[e], da, t);
            }
            t = warp_sum(t);
            for (int k = lane; k < deg; k += 32) {
                const int e = e0 + k;
                const float dm = ea[e] * (dal[e] - t);
                dal[e] = uni ? 0.f : (es[e] > 0.f ? dm : dm * kLeakySlope);      // ds
            }
        }
    }
    consumer_sync();

    if (gat) {
        // vanilla GAT: the score of edge (i, j) is s1_j + s2_i, so its gradient goes to two per-node scalars
        for (int node = tid; node < n; node += kSparseConsumers) {
            float d1 = 0.f, d2 = 0.f;
            for (int ce = colptr[node]; ce < colptr[node + 1]; ++ce) d1 += dal[cedge[ce]];
            for (int e = rowptr[node]; e < rowptr[node + 1]; ++e) d2 += dal[e];
            *reinterpret_cast<float2*>(p.gat_ds12 + ((size_t)b * n + node) * 2) = make_float2(d1, d2);
        }
        return;
    }

    // ---------------------------------------------------------------------- pass C: dK2 (rows), dU (columns), da
    for (int l = g.nch; l < n_loads; ++l) {
        const int buf = l % kSbwdBufs;
        const int c0 = (l - g.nch) * kSbwdDc;
        const int wq = min(kSbwdDc, D - c0) >> 2;
        const uint32_t uoff = (uint32_t)(buf * unit_floats) * 4u, koff = uoff + (uint32_t)g.tile_floats * 4u;
        mbar_wait(&full[buf], (uint32_t)(l / kSbwdBufs) & 1u);
        float4 da_tot = make_float4(0.f, 0.f, 0.f, 0.f), du_tot = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q8 < wq) {
            const float4 av = *reinterpret_cast<const float4*>(a_s + c0 + 4 * q8);
            float4 da_acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = warp * 4 + grp; i < n; i += kNodeStep) {
                const float4 k2 = *reinterpret_cast<const float4*>(smem_raw + swz(koff, i, q8));
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
                    const float ds = dal[e];
                    const float4 u = *reinterpret_cast<const float4*>(smem_raw + swz(uoff, meta[e] & 255u, q8));
                    const float x0 = __fadd_rn(u.x, k2.x), x1 = __fadd_rn(u.y, k2.y), x2 = __fadd_rn(u.z, k2.z), x3 = __fadd_rn(u.w, k2.w);
                    if (x0 > 0.f) { acc.x += ds; da_acc.x = fmaf(ds, x0, da_acc.x); }
                    if (x1 > 0.f) { acc.y += ds; da_acc.y = fmaf(ds, x1, da_acc.y); }
                    if (x2 > 0.f) { acc.z += ds; da_acc.z = fmaf(ds, x2, da_acc.z); }
                    if (x3 > 0.f) { acc.w += ds; da_acc.w = fmaf(ds, x3, da_acc.w); }
                }
                *reinterpret_cast<float4*>(p.dP + ((size_t)b * n + i) * p.lddp + 2 * D + c0 + 4 * q8) =
                    make_float4(av.x * acc.x, av.y * acc.y, av.z * acc.z, av.w * acc.w);
            }
            da_tot = da_acc;
            for (int j = warp * 4 + grp; j < n; j += kNodeStep) {
                const float4 u = *reinterpret_cast<const float4*>(smem_raw + swz(uoff, j, q8));
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int ce = colptr[j]; ce < colptr[j + 1]; ++ce) {
                    const int e = cedge[ce];
                    const float ds = dal[e];
                    const float4 k2 = *reinterpret_cast<const float4*>(smem_raw + swz(koff, meta[e] >> 8, q8));
                    if (__fadd_rn(u.x, k2.x) > 0.f) acc.x += ds;
                    if (__fadd_rn(u.y, k2.y) > 0.f) acc.y += ds;
                    if (__fadd_rn(u.z, k2.z) > 0.f) acc.z += ds;
                    if (__fadd_rn(u.w, k2.w) > 0.f) acc.w += ds;
                }
                const float4 du = make_float4(av.x * acc.x, av.y * acc.y, av.z * acc.z, av.w * acc.w);
                *reinterpret_cast<float4*>(p.dP + ((size_t)b * n + j) * p.lddp + D + c0 + 4 * q8) = du;
                du_tot.x += du.x; du_tot.y += du.y; du_tot.z += du.z; du_tot.w += du.w;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[buf]);
        park(p.da_partial, c0, wq, da_tot);
        if (p.du_sum != nullptr) park(p.du_sum, c0, wq, du_tot);
    }
}

inline void sparse_bwd_geometry(int n, int D, SparseBwdGeom* g) {
    g->nch = (D + kSbwdDc - 1) / kSbwdDc;
    g->tile_floats = ((n * kSbwdDc * 4 + 1023) / 1024) * 1024 / 4;
    const size_t rest = (size_t)2 * kSbwdMaxBufs * 8 + (size_t)D * 4 + (size_t)2 * n * n * 4 + (size_t)2 * (n + 1) * 4 +
                        (size_t)2 * n * n * 2 + (size_t)n + 64;
    const size_t unit = (size_t)2 * g->tile_floats * 4;
    constexpr size_t kTwoPerSm = 113 * 1024;                       // two CTAs per SM (228 KB, 1 KB reserved per CTA)
    g->bufs = (3 * unit + rest <= kTwoPerSm || 2 * unit + rest > kTwoPerSm) ? 3 : 2;
    g->smem = g->bufs * unit + rest;
}

inline int launch_graph_layer_bwd_csr(const float* P, int ldp, const float* a, const uint16_t* rowptr, const uint16_t* meta,
                                      const uint16_t* colptr, const uint16_t* cedge, const float* e_score, const float* e_alpha,
                                      const uint8_t* drop_keep, float drop_scale, const float* G, const uint8_t* relu_mask,
                                      float* dP, int lddp, float* da_partial, float* dh_sum, float* du_sum, int B, int n, int D,
                                      cudaStream_t st, float* gat_ds12 = nullptr) {
    if (B == 0) return DIGAT_OK;
    const bool gat = gat_ds12 != nullptr;
    const int pc = gat ? D : 3 * D;                                   // columns of P / dP: h only for a vanilla-GAT layer
    DIGAT_REQUIRE(P && rowptr && meta && colptr && cedge && e_score && e_alpha && G && dP && (gat || (a && da_partial)),
                  "digat_graph_layer_bwd_csr: null pointer");
    DIGAT_REQUIRE(B >= 0 && n >= 1 && n <= kPairMaxNodes, "digat_graph_layer_bwd_csr: n=%d outside [1,%d]", n, kPairMaxNodes);
    DIGAT_REQUIRE(D >= 4 && (D & 3) == 0 && D <= 1024, "digat_graph_layer_bwd_csr: D=%d must be a multiple of 4 in [4,1024]", D);
    DIGAT_REQUIRE((ldp & 3) == 0 && ldp >= pc && (lddp & 3) == 0 && lddp >= pc,
                  "digat_graph_layer_bwd_csr: ldp / lddp must be multiples of 4 and >= 3D (D for a vanilla-GAT layer)");
    DIGAT_REQUIRE(aligned16(P) && (gat || aligned16(a)) && aligned16(G) && aligned16(dP) && (!gat || (reinterpret_cast<uintptr_t>(gat_ds12) & 7) == 0),
                  "digat_graph_layer_bwd_csr: pointers must be 16-byte aligned");
    SparseBwdGeom g;
    sparse_bwd_geometry(n, D, &g);
    const DeviceInfo* di = device_info();
    if (!di) return fail(DIGAT_E_CUDA, "digat_graph_layer_bwd_csr: no CUDA device");
    if (g.smem > (size_t)di->max_smem_optin)
        return fail(DIGAT_E_UNSUPPORTED, "digat_graph_layer_bwd_csr: a graph of %d nodes needs %zu B shared memory (use the dense "
                    "digat_graph_layer_bwd)", n, g.smem);
    CUtensorMap mapG, mapP;
    int rc;
    if ((rc = make_tensor_map_2d(&mapG, G, (int64_t)B * n, D, D, n, kSbwdDc, CU_TENSOR_MAP_SWIZZLE_128B)) != DIGAT_OK) return rc;
    if ((rc = make_tensor_map_2d(&mapP, P, (int64_t)B * n, pc, ldp, n, kSbwdDc, CU_TENSOR_MAP_SWIZZLE_128B)) != DIGAT_OK) return rc;
    DIGAT_REQUIRE((gat || aligned16(da_partial)) && (!dh_sum || aligned16(dh_sum)) && (!du_sum || aligned16(du_sum)),
                  "digat_graph_layer_bwd_csr: da_partial / dh_sum / du_sum must be 16-byte aligned");
    SparseBwdArgs args{P, ldp, a, G, relu_mask, dh_sum, du_sum, gat_ds12, rowptr, meta, colptr, cedge, e_score, e_alpha, drop_keep,
                       drop_scale, dP, lddp, da_partial, B, n, D};
    if (int rc_ = ensure_dynamic_smem(graph_layer_bwd_sparse_kernel, g.smem)) return rc_;
    graph_layer_bwd_sparse_kernel<<<B, kSparseThreads, g.smem, st>>>(mapG, mapP, args, g);
    return check_launch("digat_graph_layer_bwd_csr");
}

// 1 when the training path may use the CSR kernels for graphs of n nodes (forward with per-edge outputs + this backward).
inline int graph_layer_csr_training_supported(int n, int D) {
    if (n < 1 || n > kPairMaxNodes || D < 4 || (D & 3) != 0 || D > 1024 || g_layer_mode == 1) return 0;
    const DeviceInfo* di = device_info();
    if (!di) return 0;
    SparseBwdGeom g;
    sparse_bwd_geometry(n, D, &g);
    return graph_layer_fwd_sparse_smem(n, D) <= (size_t)di->max_smem_optin && g.smem <= (size_t)di->max_smem_optin ? 1 : 0;
}

}  // namespace digat
