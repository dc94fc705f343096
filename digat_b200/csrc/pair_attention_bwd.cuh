// Backward of the fused Eq. (8) graph-attention layer (autograd of reference graphEncoders.py:150-153 / 170-173).
//
// Forward (pair_attention.cuh):  x_ijd = U_jd + K2_id,  s_ij = sum_d a_d relu(x_ijd),  e = leaky_relu(s),
//   alpha = softmax_j(mask(e)),  alpha~ = alpha * keep * scale,  Z_i = sum_j alpha~_ij h_j,  Y = relu(Z) + X.
// Given G = dY * 1[Z > 0] (the caller applies the saved relu mask) this kernel produces
//   dh_j   = sum_i alpha~_ij G_i
//   dalpha~_ij = G_i . h_j ;  dalpha = dalpha~ * keep * scale ;  dm_ij = alpha_ij (dalpha_ij - sum_k alpha_ik dalpha_ik)
//   ds_ij  = adj_ij ? dm_ij * (s_ij > 0 ? 1 : 0.2) : 0
//   dU_jd  = a_d sum_i ds_ij 1[x_ijd > 0] ;  dK2_id = a_d sum_j ds_ij 1[x_ijd > 0] ;  da_d = sum_ij ds_ij relu(x_ijd)
// The relu mask of Eq. (8) is RECOMPUTED from U and K2 (the reference's autograd stores the [B,n,n,D] tensor instead).
// One CTA per graph; two streaming passes over the feature dimension through the same 2-deep TMA pipeline as the
// forward: pass 1 streams (G, h) tiles, pass 2 streams (U, K2) tiles.  All reductions are deterministic except the
// per-CTA shared-memory accumulation of da (float atomics inside one CTA).
#pragma once
#include "common.cuh"
#include "tma.cuh"
#include "pair_attention.cuh"

namespace digat {

struct PairBwdGeom {
    int nt;            // ceil(n/4)
    int dc;            // feature chunk
    int nch;           // chunks = ceil(D/dc)
    int ldm;           // leading dim of the two n x n shared matrices (multiple of 4, >= 8*ceil(n/8))
    int tile_floats;   // floats of one [n][dc] tile rounded up to 128 bytes
    size_t smem;
};

struct PairBwdArgs {
    const float* a; const uint8_t* adj; const float* score; const float* alpha;
    const uint8_t* drop_keep; float drop_scale;
    float* dP; int lddp;
    float* da_partial;
    int B, n, D;
};

template <bool kSingleTile>
__global__ void __launch_bounds__(kPairThreads, 2)
graph_layer_bwd_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapG,
                       PairBwdArgs p, PairBwdGeom g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n, D = p.D, b = blockIdx.x;
    const int half = 2 * g.tile_floats;
    float* buf0 = reinterpret_cast<float*>(smem_raw);         // [2][half]
    float* a_s = buf0 + 2 * half;                              // [D]
    float* da_s = a_s + D;                                     // [D]
    float* M1 = da_s + D;                                      // [n][ldm]  alpha~ (row-major), later ds transposed
    float* M2 = M1 + (size_t)n * g.ldm;                        // [n][ldm]  dalpha~ / ds (row-major)
    uint64_t* full = reinterpret_cast<uint64_t*>(M2 + (size_t)n * g.ldm);

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < D; i += kPairThreads) { a_s[i] = p.a[i]; da_s[i] = 0.f; }
    // alpha~ (row-major) with zero padding
    for (int it = tid; it < n * g.ldm; it += kPairThreads) {
        const int i = it / g.ldm, j = it - i * g.ldm;
        float v = 0.f;
        if (j < n) {
            const size_t o = ((size_t)b * n + i) * n + j;
            v = p.alpha[o];
            if (p.drop_keep != nullptr) v = p.drop_keep[o] != 0 ? v * p.drop_scale : 0.f;
        }
        M1[it] = v;
    }
    __syncthreads();

    const int n_loads = 2 * g.nch;
    auto issue = [&](int l) {
        float* dst = buf0 + (l & 1) * half;
        mbar_arrive_expect_tx(&full[l & 1], 2u * n * g.dc * 4u);
        if (l < g.nch) {
            tma_load_2d(dst, &mapG, &full[l & 1], l * g.dc, b * n);                                   // G
            tma_load_2d(dst + g.tile_floats, &mapP, &full[l & 1], l * g.dc, b * n);                   // h
        } else {
            tma_load_2d(dst, &mapP, &full[l & 1], D + (l - g.nch) * g.dc, b * n);                     // U
            tma_load_2d(dst + g.tile_floats, &mapP, &full[l & 1], 2 * D + (l - g.nch) * g.dc, b * n); // K2
        }
    };
    if (tid == 0) {
        issue(0);
        if (n_loads > 1) issue(1);
    }

    // ------------------------------------------------------------------ pass 1: dalpha~ (pair tiles) and dh
    const int nt = g.nt, tiles = nt * nt;
    const int nit8 = (n + 7) >> 3;
    uint64_t acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = 0ull;

    for (int l = 0; l < g.nch; ++l) {
        const int c0 = l * g.dc;
        const int wq = min(g.dc, D - c0) >> 2;
        const uint32_t boff = (uint32_t)((l & 1) * half) * 4u;                 // G tile; h tile follows
        const uint32_t hoff = boff + (uint32_t)g.tile_floats * 4u;
        mbar_wait(&full[l & 1], (uint32_t)(l >> 1) & 1u);
        for (int t = tid; t < tiles; t += kPairThreads) {
            const int ti = t / nt, tj = t - ti * nt;
            uint32_t go[4], ho[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                go[x] = boff + (uint32_t)(min(ti + nt * x, n - 1) * g.dc) * 4u;
                ho[x] = hoff + (uint32_t)(min(tj + nt * x, n - 1) * g.dc) * 4u;
            }
            if (!kSingleTile) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = 0ull;
            }
#pragma unroll 1
            for (int q = 0; q < wq; ++q) {
                const uint32_t qo = (uint32_t)q * 16u;
                uint64_t h01[4], h23[4];
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const float4 h = *reinterpret_cast<const float4*>(smem_raw + ho[y] + qo);
                    h01[y] = pack2(h.x, h.y);
                    h23[y] = pack2(h.z, h.w);
                }
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const float4 gq = *reinterpret_cast<const float4*>(smem_raw + go[x] + qo);
                    const uint64_t g01 = pack2(gq.x, gq.y), g23 = pack2(gq.z, gq.w);
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = fma2(g01, h01[y], acc[x][y]);
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = fma2(g23, h23[y], acc[x][y]);
                }
            }
            if (!kSingleTile) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const int i = ti + nt * x, j = tj + nt * y;
                        if (i < n && j < n) {
                            float e, o;
                            unpack2(acc[x][y], e, o);
                            float* dst = M2 + i * g.ldm + j;
                            *dst = (l == 0 ? 0.f : *dst) + (e + o);
                        }
                    }
            }
        }
        // dh_j = sum_i alpha~_ij G_i : a thread owns 8 neighbour rows j x one feature quad
        const int items = nit8 * wq;
        for (int it = tid; it < items; it += kPairThreads) {
            const int q = it % wq, jb = it / wq;
            const int j0 = jb * 8;
            const float* Gs = reinterpret_cast<const float*>(smem_raw + boff) + 4 * q;
            const float* A = M1 + j0;
            uint64_t o[8][2];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k][0] = o[k][1] = 0ull;
#pragma unroll 2
            for (int i = 0; i < n; ++i) {
                const float4 gq = *reinterpret_cast<const float4*>(Gs);
                const float4 al0 = *reinterpret_cast<const float4*>(A);
                const float4 al1 = *reinterpret_cast<const float4*>(A + 4);
                Gs += g.dc;
                A += g.ldm;
                const uint64_t g01 = pack2(gq.x, gq.y), g23 = pack2(gq.z, gq.w);
                const float al[8] = {al0.x, al0.y, al0.z, al0.w, al1.x, al1.y, al1.z, al1.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t aa = pack2(al[k], al[k]);
                    o[k][0] = fma2(aa, g01, o[k][0]);
                    o[k][1] = fma2(aa, g23, o[k][1]);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int j = j0 + k;
                if (j < n) {
                    float4 y;
                    unpack2(o[k][0], y.x, y.y);
                    unpack2(o[k][1], y.z, y.w);
                    stg_stream(reinterpret_cast<float4*>(p.dP + ((size_t)b * n + j) * p.lddp + c0 + 4 * q), y);
                }
            }
        }
        __syncthreads();
        if (tid == 0 && l + 2 < n_loads) issue(l + 2);
    }
    if (kSingleTile && tid < tiles) {
        const int ti = tid / nt, tj = tid - ti * nt;
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const int i = ti + nt * x, j = tj + nt * y;
                if (i < n && j < n) {
                    float e, o;
                    unpack2(acc[x][y], e, o);
                    M2[i * g.ldm + j] = e + o;
                }
            }
    }
    __syncthreads();

    // ------------------------------------------------------------------ softmax / leaky-relu / mask backward
    for (int i = warp; i < n; i += kPairThreads / 32) {
        const size_t rowo = ((size_t)b * n + i) * n;
        float al[kPairMaxNodes / 32], da[kPairMaxNodes / 32];
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < kPairMaxNodes / 32; ++k) {
            const int j = lane + 32 * k;
            al[k] = 0.f; da[k] = 0.f;
            if (j < n) {
                al[k] = p.alpha[rowo + j];
                float d = M2[i * g.ldm + j];
                if (p.drop_keep != nullptr) d = p.drop_keep[rowo + j] != 0 ? d * p.drop_scale : 0.f;
                da[k] = d;
                t = fmaf(al[k], d, t);
            }
        }
        t = warp_sum(t);
#pragma unroll
        for (int k = 0; k < kPairMaxNodes / 32; ++k) {
            const int j = lane + 32 * k;
            if (j < n) {
                const float dm = al[k] * (da[k] - t);
                const float slope = p.score[rowo + j] > 0.f ? 1.f : kLeakySlope;
                const float ds = p.adj[rowo + j] != 0 ? dm * slope : 0.f;
                M2[i * g.ldm + j] = ds;          // row-major  ds[i][j]
                M1[j * g.ldm + i] = ds;          // transposed ds[j][i]
            }
        }
    }
    __syncthreads();
    for (int it = tid; it < n * (g.ldm - n); it += kPairThreads) {      // zero the padding columns of both matrices
        const int pad = g.ldm - n;
        const int r = it / pad, c = n + it - r * pad;
        M1[r * g.ldm + c] = 0.f;
        M2[r * g.ldm + c] = 0.f;
    }
    __syncthreads();

    // ------------------------------------------------------------------ pass 2: Eq. (8) backward (relu mask recomputed)
    for (int l = g.nch; l < n_loads; ++l) {
        const int c0 = (l - g.nch) * g.dc;
        const int wq = min(g.dc, D - c0) >> 2;
        const float* Us = buf0 + (l & 1) * half;
        const float* K2s = Us + g.tile_floats;
        mbar_wait(&full[l & 1], (uint32_t)(l >> 1) & 1u);
        const int items = 2 * nt * wq;
        for (int it = tid; it < items; it += kPairThreads) {
            const bool passB = it >= nt * wq;                  // A: owns query rows i (dK2, da);  B: owns neighbour rows j (dU)
            const int w = passB ? it - nt * wq : it;
            const int q = w % wq, rb = w / wq;
            const int r0 = rb * 4;
            const float* own = (passB ? Us : K2s) + 4 * q;     // the 4 rows this thread owns
            const float* oth = (passB ? K2s : Us) + 4 * q;     // the rows it loops over
            const float* W = (passB ? M2 : M1) + r0;           // ds[other][own 0..3]
            float4 ow[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) ow[x] = *reinterpret_cast<const float4*>(own + min(r0 + x, n - 1) * g.dc);
            float4 accm[4];
            float4 dacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int x = 0; x < 4; ++x) accm[x] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
            for (int o = 0; o < n; ++o) {
                const float4 ot = *reinterpret_cast<const float4*>(oth);
                const float4 dsv = *reinterpret_cast<const float4*>(W);
                oth += g.dc;
                W += g.ldm;
                const float dsx[4] = {dsv.x, dsv.y, dsv.z, dsv.w};
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    // same operands and the same IEEE add as the forward: the recomputed mask matches its relu exactly
                    const float x0 = ot.x + ow[x].x, x1 = ot.y + ow[x].y, x2 = ot.z + ow[x].z, x3 = ot.w + ow[x].w;
                    accm[x].x += x0 > 0.f ? dsx[x] : 0.f;
                    accm[x].y += x1 > 0.f ? dsx[x] : 0.f;
                    accm[x].z += x2 > 0.f ? dsx[x] : 0.f;
                    accm[x].w += x3 > 0.f ? dsx[x] : 0.f;
                    if (!passB) {
                        dacc.x = fmaf(dsx[x], fmaxf(x0, 0.f), dacc.x);
                        dacc.y = fmaf(dsx[x], fmaxf(x1, 0.f), dacc.y);
                        dacc.z = fmaf(dsx[x], fmaxf(x2, 0.f), dacc.z);
                        dacc.w = fmaf(dsx[x], fmaxf(x3, 0.f), dacc.w);
                    }
                }
            }
            const float4 av = *reinterpret_cast<const float4*>(a_s + c0 + 4 * q);
            const int colbase = (passB ? D : 2 * D) + c0 + 4 * q;
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int r = r0 + x;
                if (r < n) {
                    float4 y;
                    y.x = av.x * accm[x].x; y.y = av.y * accm[x].y; y.z = av.z * accm[x].z; y.w = av.w * accm[x].w;
                    stg_stream(reinterpret_cast<float4*>(p.dP + ((size_t)b * n + r) * p.lddp + colbase), y);
                }
            }
            if (!passB) {
                // rows r0+x >= n were clamped to row n-1 for the loads, but their ds column is zero padding: no effect
                atomicAdd(da_s + c0 + 4 * q + 0, dacc.x);
                atomicAdd(da_s + c0 + 4 * q + 1, dacc.y);
                atomicAdd(da_s + c0 + 4 * q + 2, dacc.z);
                atomicAdd(da_s + c0 + 4 * q + 3, dacc.w);
            }
        }
        __syncthreads();
        if (tid == 0 && l + 2 < n_loads) issue(l + 2);
    }
    for (int i = tid; i < D; i += kPairThreads) p.da_partial[(size_t)b * D + i] = da_s[i];
}

inline void pair_bwd_geometry(int n, int D, PairBwdGeom* g) {
    g->nt = (n + 3) / 4;
    g->ldm = ((n + 7) / 8) * 8 + 4;
    // Feature chunk: 68 (68/4 = 17 odd: conflict-free row interleave), shrunk in steps of 8 (dc/4 stays odd) until TWO CTAs fit
    // an SM.  At n = 68, dc = 68 the kernel needed 118.8 KB: one CTA per SM, so a training batch of 320 graphs ran in three
    // waves on 148 SMs (ncu: 670 us); with dc = 60 (109.6 KB) the 296 slots take it in one wave plus a short tail.
    const size_t two_per_sm = (size_t)113 * 1024;
    int dc = 68;
    if (dc > D) dc = D;
    for (;;) {
        g->dc = dc;
        g->nch = (D + dc - 1) / dc;
        g->tile_floats = ((n * dc * 4 + 127) / 128) * 128 / 4;
        g->smem = (size_t)4 * g->tile_floats * 4 + (size_t)2 * D * 4 + (size_t)2 * n * g->ldm * 4 + 16;
        if (g->smem <= two_per_sm || dc <= 36 || dc >= D) break;
        dc -= 8;
    }
}

inline int launch_graph_layer_bwd(const float* P, int ldp, const float* a, const uint8_t* adj, const float* score,
                                  const float* alpha, const uint8_t* drop_keep, float drop_scale, const float* G,
                                  float* dP, int lddp, float* da_partial, int B, int n, int D, cudaStream_t st) {
    if (B == 0) return DIGAT_OK;
    DIGAT_REQUIRE(P && a && adj && score && alpha && G && dP && da_partial, "digat_graph_layer_bwd: null pointer");
    DIGAT_REQUIRE(B >= 0 && n >= 1 && n <= kPairMaxNodes, "digat_graph_layer_bwd: n=%d outside [1,%d]", n, kPairMaxNodes);
    DIGAT_REQUIRE(D >= 4 && (D & 3) == 0 && D <= 1024, "digat_graph_layer_bwd: D=%d must be a multiple of 4 in [4,1024]", D);
    DIGAT_REQUIRE((ldp & 3) == 0 && ldp >= 3 * D && (lddp & 3) == 0 && lddp >= 3 * D,
                  "digat_graph_layer_bwd: ldp / lddp must be multiples of 4 and >= 3D");
    DIGAT_REQUIRE(aligned16(P) && aligned16(a) && aligned16(G) && aligned16(dP),
                  "digat_graph_layer_bwd: pointers must be 16-byte aligned");
    if (B == 0) return DIGAT_OK;
    PairBwdGeom g;
    pair_bwd_geometry(n, D, &g);
    const DeviceInfo* di = device_info();
    if (!di) return fail(DIGAT_E_CUDA, "digat_graph_layer_bwd: no CUDA device");
    DIGAT_REQUIRE(g.smem <= (size_t)di->max_smem_optin, "digat_graph_layer_bwd: needs %zu B shared memory", g.smem);
    CUtensorMap mapP, mapG;
    int rc;
    if ((rc = make_tensor_map_2d(&mapP, P, (int64_t)B * n, 3 * D, ldp, n, g.dc, CU_TENSOR_MAP_SWIZZLE_NONE)) != DIGAT_OK) return rc;
    if ((rc = make_tensor_map_2d(&mapG, G, (int64_t)B * n, D, D, n, g.dc, CU_TENSOR_MAP_SWIZZLE_NONE)) != DIGAT_OK) return rc;
    PairBwdArgs args{a, adj, score, alpha, drop_keep, drop_scale, dP, lddp, da_partial, B, n, D};
    const bool single = g.nt * g.nt <= kPairThreads;
    if (single) {
        if (int rc_ = ensure_dynamic_smem(graph_layer_bwd_kernel<true>, (size_t)(g.smem))) return rc_;
        graph_layer_bwd_kernel<true><<<B, kPairThreads, g.smem, st>>>(mapP, mapG, args, g);
    } else {
        if (int rc_ = ensure_dynamic_smem(graph_layer_bwd_kernel<false>, (size_t)(g.smem))) return rc_;
        graph_layer_bwd_kernel<false><<<B, kPairThreads, g.smem, st>>>(mapP, mapG, args, g);
    }
    return check_launch("digat_graph_layer_bwd");
}

}  // namespace digat
