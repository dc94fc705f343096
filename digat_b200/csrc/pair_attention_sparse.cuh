// Fused Eq. (8) graph-attention layer, forward, EDGE-DRIVEN variant (inference).
//
// The reference evaluates s_ij for all n^2 pairs and then overwrites the non-edges with -1e9
// (graphEncoders.py:150-152).  A masked pair contributes exp(-1e9 - max) == 0.0f to the softmax of any row that has at
// least one edge, so its score is never observable: this kernel computes s_ij ONLY on the edges of the adjacency and
// aggregates over neighbours only -- the same alpha and Y as the dense evaluation, with E instead of n^2 pair
// evaluations (MIND-shaped user graphs: E ~ 300-650 of 4624; SAG trees: ~2n of n^2).  A row without any edge (cannot
// happen in the reference's data: the diagonal is always set) is handled like the reference: every entry is -1e9, the
// softmax is uniform 1/n over ALL nodes of its graph.
//
// A CTA owns one graph, or G small graphs treated as one block-diagonal graph of G*n <= 128 nodes (kMulti; the rows
// of consecutive graphs are contiguous in P / X / Y / adj, so one TMA box covers them).  The kernel is organised around
// the loads:
//   producer  (last warp) walks the load schedule (13 U|K2 units of 32 features, then 7 h units of 64 features) through
//             a ring of shared-memory buffers, waiting only on per-buffer "empty" mbarriers;
//   consumers no CTA-wide barrier inside the streaming loops -- a warp waits on "full[buf]", does its share and arrives
//             on "empty[buf]"; fast warps run up to a ring's depth ahead of slow ones.
//   setup     adjacency bytes -> CSR (row pointers, one uint16 (column | row << 8) per edge) with warp ballots + one warp
//             scan; nodes pruned by row_active get no CSR row
//   phase 1   one EDGE per thread: U_j and K2_i quads from smem, packed FADD2 / FMNMX / FFMA2, score[e] += partial
//   phase 2   one warp per row: leaky-relu, max/sum by shuffles over the row's CSR segment, exp, normalise
//   phase 3   half a warp per row, one lane per feature quad: loop over the row's neighbours (FFMA2, four neighbours in
//             flight), relu + residual; pruned rows are neither read nor written; with Yc the row is also written to the
//             compact operand of the next layer's projection
// In indexed (de-duplicated) mode the shared K1 tile becomes this pair's U in registers: fl(fl(K1 + k3) + K2), the
// reference's rounding.  HBM traffic per graph: 5nD*4 + n^2 bytes (pruned rows of P are still streamed by the TMA boxes).
#pragma once
#include "common.cuh"
#include "tma.cuh"
#include "pair_attention.cuh"

namespace digat {

#ifndef DIGAT_SPARSE_WARPS
#define DIGAT_SPARSE_WARPS 10
#endif
#ifndef DIGAT_SPARSE_BUFS
#define DIGAT_SPARSE_BUFS 4
#endif
#ifndef DIGAT_SPARSE_MINCTAS
#define DIGAT_SPARSE_MINCTAS 2
#endif
#ifndef DIGAT_SPARSE_UNROLL
#define DIGAT_SPARSE_UNROLL 4      // measured (tools/run_layer_variants.sh): 4 quads in flight 0.656 ms vs 0.683 (2); more warps are slower
#endif
constexpr int kSparseConsumers = 32 * DIGAT_SPARSE_WARPS; // consumer warps 0..W-1
constexpr int kSparseThreads = kSparseConsumers + 32;     // + producer warp
constexpr int kSparseBufs = DIGAT_SPARSE_BUFS;
constexpr int kSparseDc1 = 32, kSparseDc3 = 64;
constexpr int kSparseUnroll = DIGAT_SPARSE_UNROLL;      // feature quads of one edge in flight in the score loop

struct SparseGeom {
    int nch1, nch3;
    int G;                 // graphs per CTA: small graphs are batched into one block-diagonal graph of G*n nodes
    int stagger_lo, stagger_hi;   // CTAs [lo, hi) (the second resident of every SM in the first wave) start late (experiment)
    int tile_floats;       // floats of one [G*n][32] tile rounded up to 1024 bytes (SWIZZLE_128B atoms stay aligned)
    int unit_floats;       // floats of one ring buffer = 2 tiles (>= the [G*n][64] h tile)
    size_t smem;
};

__device__ __forceinline__ void consumer_sync() {         // named barrier over the 10 consumer warps only
    asm volatile("bar.sync 1, %0;" :: "n"(kSparseConsumers) : "memory");
}

// kMulti: several small graphs per CTA (block-diagonal); false keeps every index a function of n alone (one graph per CTA)
// kTrain: the training extras (per-edge score / alpha, attention dropout, relu mask) are compiled in only for the
// instantiation the training path launches -- with them in the inference kernels the layer ran 5 % slower.
template <bool kIndexed, bool kMulti, bool kTrain = false>
__global__ void __launch_bounds__(kSparseThreads, DIGAT_SPARSE_MINCTAS)
graph_layer_fwd_sparse_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map3,
                              PairAttnArgs p, SparseGeom g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n, D = p.D;
    const int b = kMulti ? blockIdx.x * g.G : blockIdx.x;          // first graph of this CTA (G == 1 in indexed mode)
    const int N = kMulti ? min(g.G, p.B - b) * n : n;              // nodes of the block-diagonal graph this CTA evaluates
    const int NB = kMulti ? g.G * n : n;                           // rows of a TMA box
    const int src0 = kIndexed ? p.px_index[b] : b;

    float* ring = reinterpret_cast<float*>(smem_raw);              // [kSparseBufs][unit_floats]
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + kSparseBufs * g.unit_floats);   // [kSparseBufs]
    uint64_t* empty = full + kSparseBufs;                          // [kSparseBufs]
    float* a_s = reinterpret_cast<float*>(empty + kSparseBufs);    // [D]
    float* k3_s = a_s + D;                                         // [D]
    float* score = k3_s + D;                                       // [NB*n] per-edge score, later alpha~
    int* rowptr = reinterpret_cast<int*>(score + NB * n);          // [NB+2]
    uint16_t* meta = reinterpret_cast<uint16_t*>(rowptr + (NB + 2)); // [NB*n] per edge: neighbour (column) | query row << 8
    uint8_t* adj_s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(meta + NB * n) + 15) & ~(uintptr_t)15);   // [NB*n], 16-byte aligned
    uint8_t* uniform_row = adj_s + ((NB * n + 15) & ~15);          // [NB] 1 = row without edges (uniform softmax)
    uint8_t* dead_row = uniform_row + NB;                          // [NB] 1 = pruned node (row_active == 0): no edges, Y = X
    int* pos_s = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(dead_row + NB) + 3) & ~(uintptr_t)3);   // [NB] compact row of each node
    int* n_act_s = pos_s + NB;                                     // number of rows this CTA evaluates (not pruned)
    uint8_t* act_list = reinterpret_cast<uint8_t*>(n_act_s + 1);   // [NB] their node indices, ascending: phase 3 walks THIS list

    if (tid == 0) {
        for (int i = 0; i < kSparseBufs; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kSparseConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#ifdef DIGAT_SPARSE_STAGGER
    // The CTAs of a wave run in lockstep (same graph sizes): everyone streams P at the same time (DRAM-bound) and everyone
    // builds its CSR / does its softmax at the same time (DRAM idle).  Delaying the second CTA of every SM once, in the first
    // wave, by about half a graph puts the two residents of an SM out of phase for the rest of the launch.
    if (blockIdx.x >= gridDim.y * 0 + (unsigned)g.stagger_lo && blockIdx.x < (unsigned)g.stagger_hi) __nanosleep(DIGAT_SPARSE_STAGGER);
#endif
    __syncthreads();
    const int n_loads = g.nch1 + g.nch3;

    if (warp == kSparseConsumers / 32) {
        // ------------------------------------------------------------------ producer warp
        if (lane == 0) {
            for (int l = 0; l < n_loads; ++l) {
                const int buf = l % kSparseBufs;
                float* dst = ring + buf * g.unit_floats;
                mbar_wait(&empty[buf], ((uint32_t)(l / kSparseBufs) & 1u) ^ 1u);
                if (l < g.nch1) {
                    mbar_arrive_expect_tx(&full[buf], 2u * NB * kSparseDc1 * 4u);       // (rows past the tensor end arrive as zeros)
                    tma_load_2d(dst, &map1, &full[buf], D + l * kSparseDc1, src0 * n);                     // U (or K1)
                    tma_load_2d(dst + g.tile_floats, &map1, &full[buf], 2 * D + l * kSparseDc1, src0 * n);  // K2
                } else {
                    mbar_arrive_expect_tx(&full[buf], (uint32_t)NB * kSparseDc3 * 4u);
                    tma_load_2d(dst, &map3, &full[buf], (l - g.nch1) * kSparseDc3, src0 * n);              // h
                }
            }
        }
        return;                                                    // consumers synchronise among themselves (barrier 1)
    }

    // ---------------------------------------------------------------------- consumers: setup (overlaps the first loads)
#ifdef DIGAT_TC_TIMING
    const long long t0 = clock64();
    long long ta = t0, wait1 = 0, wait3 = 0;
#endif
    const float* __restrict__ gat_s = p.gat_s;                     // vanilla-GAT scores: no phase-1 loads (g.nch1 == 0)
    for (int i = tid; i < D / 4; i += kSparseConsumers) {
        if (gat_s == nullptr) reinterpret_cast<float4*>(a_s)[i] = reinterpret_cast<const float4*>(p.a)[i];
        if (kIndexed) reinterpret_cast<float4*>(k3_s)[i] = reinterpret_cast<const float4*>(p.k3 + (size_t)b * p.ldk3)[i];
    }
    const bool have_csr = !kMulti && p.csr_rowptr != nullptr;
    if (have_csr) {
        // CSR built once per batch and graph (digat_build_graph_csr): two short loads instead of the ballot passes below
        const size_t cg = p.csr_index != nullptr ? (size_t)p.csr_index[b] : (size_t)b;
        const uint16_t* rp = p.csr_rowptr + cg * (n + 1);
        for (int i = tid; i <= n; i += kSparseConsumers) {
            const uint32_t raw = rp[i];
            rowptr[i] = (int)(raw & 0x7fffu);
            if (i > 0) uniform_row[i - 1] = (uint8_t)(raw >> 15);
        }
        if (p.Yc != nullptr)
            for (int i = tid; i < N; i += kSparseConsumers) pos_s[i] = p.row_pos[(size_t)b * n + i];
        consumer_sync();
        const int Ec = rowptr[n];
        const uint16_t* mg = p.csr_meta + cg * (size_t)n * n;              // (n*n*2 bytes per record: 8-byte aligned for even n)
        for (int e = tid; e < Ec; e += kSparseConsumers) {
            const uint32_t mt = mg[e];
            meta[e] = (uint16_t)mt;
            score[e] = gat_s == nullptr ? 0.f
                                        : gat_s[((size_t)b * n + (mt & 255u)) * 2] + gat_s[((size_t)b * n + (mt >> 8)) * 2 + 1];
        }
        for (int i = tid; i < N; i += kSparseConsumers) dead_row[i] = rowptr[i + 1] == rowptr[i];
        if (warp == 0) {
            int run_a = 0;
            for (int base = 0; base < N; base += 32) {
                const int i = base + lane;
                const bool alive = i < N && rowptr[i + 1] != rowptr[i];
                const unsigned m = __ballot_sync(0xffffffffu, alive);
                if (alive) act_list[run_a + __popc(m & ((1u << lane) - 1u))] = (uint8_t)i;
                run_a += __popc(m);
            }
            if (lane == 0) *n_act_s = run_a;
        }
    } else {
    {
        // adjacency -> shared memory with independent 16-byte loads (one round trip) when the graph is 16-byte aligned
        // (the graphs of one CTA are consecutive in memory: several graphs per CTA only without adj_index)
        const uint8_t* adj_g = p.adj + (p.adj_index != nullptr ? (size_t)p.adj_index[b] : (size_t)b) * n * n;
        const int n2 = N * n;
        if (((n2 | (int)(reinterpret_cast<uintptr_t>(adj_g) & 15)) & 15) == 0) {
            for (int i = tid; i < n2 / 16; i += kSparseConsumers)
                reinterpret_cast<uint4*>(adj_s)[i] = __ldg(reinterpret_cast<const uint4*>(adj_g) + i);
        } else {
            for (int i = tid; i < n2; i += kSparseConsumers) adj_s[i] = adj_g[i];
        }
        for (int i = tid; i < N; i += kSparseConsumers) {
            dead_row[i] = (p.row_active != nullptr && p.row_active[(size_t)b * n + i] == 0) ? 1 : 0;
            if (p.Yc != nullptr) pos_s[i] = p.row_pos[(size_t)b * n + i];
        }
    }
    consumer_sync();
#ifdef DIGAT_TC_TIMING
    ta = clock64();
#endif
    // CSR pass A: degrees (a row without edges becomes a full row with uniform weights; a pruned row has no edges)
    for (int i = warp; i < N; i += kSparseConsumers / 32) {        // row i of the block-diagonal graph = adjacency row i, n columns
        const bool dead = dead_row[i] != 0;
        int deg = 0;
        if (!dead)
            for (int k = 0; k < (n + 31) / 32; ++k) {
                const int j = lane + 32 * k;
                deg += __popc(__ballot_sync(0xffffffffu, j < n && adj_s[i * n + j] != 0));
            }
        if (lane == 0) {
            uniform_row[i] = deg == 0 && !dead;
            rowptr[i + 1] = dead ? 0 : (deg == 0 ? n : deg);       // degrees for now, scanned below
        }
    }
    consumer_sync();
    if (warp == 0) {                                               // inclusive scan of <= 128 entries by one warp
        int run_e = 0;
        for (int base = 0; base < N; base += 32) {
            const int i = base + lane;
            int ve = i < N ? rowptr[i + 1] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int te = __shfl_up_sync(0xffffffffu, ve, o);
                if (lane >= o) ve += te;
            }
            if (i < N) rowptr[i + 1] = run_e + ve;
            run_e += __shfl_sync(0xffffffffu, ve, 31);
        }
        if (lane == 0) rowptr[0] = 0;
        int run_a = 0;                                             // compact list of the evaluated rows (phase 3 assigns half
        for (int base = 0; base < N; base += 32) {                 // warps to list entries, so pruned rows cost no pass)
            const int i = base + lane;
            const bool alive = i < N && dead_row[i] == 0;
            const unsigned m = __ballot_sync(0xffffffffu, alive);
            if (alive) act_list[run_a + __popc(m & ((1u << lane) - 1u))] = (uint8_t)i;
            run_a += __popc(m);
        }
        if (lane == 0) *n_act_s = run_a;
    }
    consumer_sync();
    // CSR pass B: column / row index of every edge, zeroed scores
    for (int i = warp; i < N; i += kSparseConsumers / 32) {
        const int e0 = rowptr[i];
        const bool uni = uniform_row[i] != 0;
        if (dead_row[i] != 0) continue;
        const int node0 = kMulti ? (i / n) * n : 0;                // first node of row i's graph inside the CTA
        int filled = 0;
        for (int k = 0; k < (n + 31) / 32; ++k) {
            const int j = lane + 32 * k;
            const bool on = j < n && (uni || adj_s[i * n + j] != 0);
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (on) {
                const int e = e0 + filled + __popc(m & ((1u << lane) - 1u));
                meta[e] = (uint16_t)((node0 + j) | (i << 8));
                // Eq. (8) scores are accumulated in phase 1; a vanilla-GAT score is complete here: fl(a1.h_j + a2.h_i)
                score[e] = gat_s == nullptr ? 0.f
                                            : gat_s[((size_t)b * n + node0 + j) * 2] + gat_s[((size_t)b * n + i) * 2 + 1];
            }
            filled += __popc(m);
        }
    }
    }
    const int E = rowptr[N];
    consumer_sync();

#ifdef DIGAT_TC_TIMING
    const long long t1 = clock64();
#endif
    // ---------------------------------------------------------------------- phase 1: scores on the edges only
    const uint32_t a_off = (uint32_t)(reinterpret_cast<uint8_t*>(a_s) - smem_raw);
    const uint32_t k3_off = (uint32_t)(reinterpret_cast<uint8_t*>(k3_s) - smem_raw);
    for (int l = 0; l < g.nch1; ++l) {
        const int buf = l % kSparseBufs;
        const int c0 = l * kSparseDc1;
        const int wq = min(kSparseDc1, D - c0) >> 2;
        const uint32_t uoff = (uint32_t)(buf * g.unit_floats) * 4u;            // the ring starts at dynamic smem offset 0
        const uint32_t koff = uoff + (uint32_t)g.tile_floats * 4u;
        const uint32_t aoff = a_off + (uint32_t)c0 * 4u, k3off = k3_off + (uint32_t)c0 * 4u;
#ifdef DIGAT_TC_TIMING
        const long long tw0 = clock64();
#endif
        mbar_wait(&full[buf], (uint32_t)(l / kSparseBufs) & 1u);
#ifdef DIGAT_TC_TIMING
        wait1 += clock64() - tw0;
#endif
        // One edge per thread; a thread that owns a second edge (E > 320) evaluates BOTH in the same pass over the feature
        // quads -- their loads are independent, so the second edge rides in the latency shadow of the first (the loop is
        // latency-bound: a second sequential pass used to double the time of every unit), and the a / k3 quads are shared.
        for (int eb = warp * 32; eb < E; eb += 2 * kSparseConsumers) {        // warp-uniform trip count (__any_sync below)
            // rows are 128 bytes (all rows would start in bank 0): the tiles are loaded with SWIZZLE_128B, i.e. the 16-byte
            // chunk q of the 128-byte line L sits at chunk q ^ (L & 7) -- threads on different rows hit different banks
            const int e = eb + lane, e2 = e + kSparseConsumers;
            const bool one = e < E, two = e2 < E;                   // (a lane past E re-evaluates edge 0 and stores nothing)
            const uint32_t mt = meta[one ? e : 0], mt2 = meta[two ? e2 : 0];
            const uint32_t uo = uoff + (mt & 255u) * (uint32_t)(kSparseDc1 * 4), ko = koff + (mt >> 8) * (uint32_t)(kSparseDc1 * 4);
            const uint32_t uo2 = uoff + (mt2 & 255u) * (uint32_t)(kSparseDc1 * 4), ko2 = koff + (mt2 >> 8) * (uint32_t)(kSparseDc1 * 4);
            const uint32_t ukey = ((uo >> 7) & 7u) << 4, kkey = ((ko >> 7) & 7u) << 4;
            const uint32_t ukey2 = ((uo2 >> 7) & 7u) << 4, kkey2 = ((ko2 >> 7) & 7u) << 4;
            uint64_t acc0 = 0ull, acc1 = 0ull, bcc0 = 0ull, bcc1 = 0ull;
            if (!__any_sync(0xffffffffu, two)) {                    // no lane of this warp owns a second edge: plain loop
#pragma unroll kSparseUnroll
                for (int q = 0; q < wq; ++q) {
                    const uint32_t qo = (uint32_t)q * 16u;
                    const float4 av = *reinterpret_cast<const float4*>(smem_raw + aoff + qo);
                    const float4 k2 = *reinterpret_cast<const float4*>(smem_raw + ko + (qo ^ kkey));
                    const float4 u = *reinterpret_cast<const float4*>(smem_raw + uo + (qo ^ ukey));
                    uint64_t u01 = pack2(u.x, u.y), u23 = pack2(u.z, u.w);
                    if (kIndexed) {
                        const float4 kk = *reinterpret_cast<const float4*>(smem_raw + k3off + qo);
                        u01 = add2(pack2(kk.x, kk.y), u01);
                        u23 = add2(pack2(kk.z, kk.w), u23);
                    }
                    float s0, s1, s2, s3;
                    unpack2(add2(u01, pack2(k2.x, k2.y)), s0, s1);
                    unpack2(add2(u23, pack2(k2.z, k2.w)), s2, s3);
                    acc0 = fma2(pack2(av.x, av.y), pack2(fmaxf(s0, 0.f), fmaxf(s1, 0.f)), acc0);
                    acc1 = fma2(pack2(av.z, av.w), pack2(fmaxf(s2, 0.f), fmaxf(s3, 0.f)), acc1);
                }
            } else {
#pragma unroll kSparseUnroll
                for (int q = 0; q < wq; ++q) {
                    const uint32_t qo = (uint32_t)q * 16u;
                    const float4 av = *reinterpret_cast<const float4*>(smem_raw + aoff + qo);
                    const float4 k2 = *reinterpret_cast<const float4*>(smem_raw + ko + (qo ^ kkey));
                    const float4 u = *reinterpret_cast<const float4*>(smem_raw + uo + (qo ^ ukey));
                    const float4 k2b = *reinterpret_cast<const float4*>(smem_raw + ko2 + (qo ^ kkey2));
                    const float4 ub = *reinterpret_cast<const float4*>(smem_raw + uo2 + (qo ^ ukey2));
                    uint64_t u01 = pack2(u.x, u.y), u23 = pack2(u.z, u.w);
                    uint64_t v01 = pack2(ub.x, ub.y), v23 = pack2(ub.z, ub.w);
                    if (kIndexed) {                                     // the staged tile is K1: U = fl(k3 + K1)
                        const float4 kk = *reinterpret_cast<const float4*>(smem_raw + k3off + qo);
                        const uint64_t k01 = pack2(kk.x, kk.y), k23 = pack2(kk.z, kk.w);
                        u01 = add2(k01, u01);
                        u23 = add2(k23, u23);
                        v01 = add2(k01, v01);
                        v23 = add2(k23, v23);
                    }
                    const uint64_t a01 = pack2(av.x, av.y), a23 = pack2(av.z, av.w);
                    float s0, s1, s2, s3;
                    unpack2(add2(u01, pack2(k2.x, k2.y)), s0, s1);
                    unpack2(add2(u23, pack2(k2.z, k2.w)), s2, s3);
                    acc0 = fma2(a01, pack2(fmaxf(s0, 0.f), fmaxf(s1, 0.f)), acc0);
                    acc1 = fma2(a23, pack2(fmaxf(s2, 0.f), fmaxf(s3, 0.f)), acc1);
                    unpack2(add2(v01, pack2(k2b.x, k2b.y)), s0, s1);
                    unpack2(add2(v23, pack2(k2b.z, k2b.w)), s2, s3);
                    bcc0 = fma2(a01, pack2(fmaxf(s0, 0.f), fmaxf(s1, 0.f)), bcc0);
                    bcc1 = fma2(a23, pack2(fmaxf(s2, 0.f), fmaxf(s3, 0.f)), bcc1);
                }
            }
            float a0, a1, b0, b1;
            unpack2(acc0, a0, a1);
            unpack2(acc1, b0, b1);
            if (one) score[e] += (a0 + a1) + (b0 + b1);            // each edge is owned by exactly one thread
            if (two) {
                unpack2(bcc0, a0, a1);
                unpack2(bcc1, b0, b1);
                score[e2] += (a0 + a1) + (b0 + b1);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[buf]);                   // this warp is done with the buffer
    }
    consumer_sync();                                               // every edge score is complete
#ifdef DIGAT_TC_TIMING
    const long long t2 = clock64();
#endif

    // ---------------------------------------------------------------------- phase 2: softmax over each row's edges
    // Training extras (one graph per CTA, precomputed CSR): the raw score s_e and the softmax weight alpha_e of every edge go
    // out in CSR order (score_out / alpha_out [B, n*n], what the edge-driven backward reads), and the dropout on the attention
    // weights (graphEncoders.py:152/172) is applied here: alpha~_e = alpha_e * keep[i,j] * scale.
    const bool train_out = kTrain && p.alpha_out != nullptr;
    float* __restrict__ e_alpha = train_out ? p.alpha_out + (size_t)b * n * n : nullptr;
    float* __restrict__ e_score = (train_out && p.score_out != nullptr) ? p.score_out + (size_t)b * n * n : nullptr;
    const uint8_t* __restrict__ keep_g = (kTrain && p.drop_keep != nullptr) ? p.drop_keep + (size_t)b * n * n : nullptr;
    auto finish_edge = [&](int e, float s_raw, float al) {           // store alpha~ for phase 3, the raw values for the backward
        if (e_score != nullptr) e_score[e] = s_raw;
        if (e_alpha != nullptr) e_alpha[e] = al;
        if (keep_g != nullptr) {
            const uint32_t mt = meta[e];
            al = keep_g[(mt >> 8) * n + (mt & 255u)] != 0 ? al * p.drop_scale : 0.f;
        }
        score[e] = al;
    };
    for (int i = warp; i < N; i += kSparseConsumers / 32) {
        const int e0 = rowptr[i], deg = rowptr[i + 1] - e0;
        if (deg == 0) continue;                                    // pruned row
        const bool uni = uniform_row[i] != 0;
        if (deg <= 32) {                                           // the common case: one edge per lane
            float m = -INFINITY, s = 0.f;
            if (lane < deg) {
                s = score[e0 + lane];
                m = uni ? kNegFill : (s > 0.f ? s : s * kLeakySlope);
            }
            const float mx1 = warp_max(m);
            const float ex = lane < deg ? expf(m - mx1) : 0.f;
            const float sum1 = warp_sum(ex);
            if (lane < deg) finish_edge(e0 + lane, s, ex / sum1);
            continue;
        }
        float v[kPairMaxNodes / 32], sr[kPairMaxNodes / 32];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < kPairMaxNodes / 32; ++k) {
            const int t = lane + 32 * k;
            float m = -INFINITY;
            sr[k] = 0.f;
            if (t < deg) {
                const float s = score[e0 + t];
                sr[k] = s;
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
            if (t < deg) finish_edge(e0 + t, sr[k], v[k] / sum);
        }
    }
    consumer_sync();

#ifdef DIGAT_TC_TIMING
    const long long t3 = clock64();
#endif
    // ---------------------------------------------------------------------- phase 3: Y = relu(sum_{j in N(i)} alpha_ij h_j) + X
    // Two rows per warp pass, 16 lanes per row.  The h units (64 features each, one ring buffer) are consumed in PAIRS: a lane
    // owns quad q of unit A and quad q of unit B, so the per-neighbour bookkeeping (alpha, neighbour index, loop control) and
    // the per-row bookkeeping (row pointers, addresses) are paid once per 128 features -- this phase was 35 % of the kernel's
    // instructions with one unit at a time.
    const int half_lane = lane & 15, sub = lane >> 4;
    // Residual rows: a pass over a row's few neighbours is far shorter than a DRAM round trip, so the x values of the next
    // unit pair (up to kXSlots passes of this warp) are requested one iteration AHEAD and sit in registers meanwhile.
    constexpr int kXSlots = 2;
    constexpr int kRowPairStep = kSparseConsumers / 32;
    const int n_act = *n_act_s;
    auto row_of = [&](int slot) {                                  // node evaluated by this half warp in pass `slot`, or -1
        const int k = 2 * (warp + slot * kRowPairStep) + sub;
        return k < n_act ? (int)act_list[k] : -1;
    };
    auto load_x = [&](int unit, int slot) {                        // quad of this lane in h unit `unit` of the row of pass `slot`
        const int c0_ = unit * kSparseDc3;
        const int i_ = row_of(slot);
        return (unit < g.nch3 && i_ >= 0 && c0_ + 4 * half_lane < D)
                   ? ldg_stream(reinterpret_cast<const float4*>(p.X + ((size_t)src0 * n + i_) * D + c0_ + 4 * half_lane))
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 xqa[kXSlots], xqb[kXSlots];
#pragma unroll
    for (int sl = 0; sl < kXSlots; ++sl) {
        xqa[sl] = load_x(0, sl);
        xqb[sl] = load_x(1, sl);
    }
    for (int l = g.nch1; l < n_loads; l += 2) {
        const bool pair = l + 1 < n_loads;                          // the last iteration may hold a single unit
        const int bufa = l % kSparseBufs, bufb = (l + 1) % kSparseBufs;
        const int unit = l - g.nch1;
        const int c0 = unit * kSparseDc3;
        const int wqa = min(kSparseDc3, D - c0) >> 2;
        const int wqb = pair ? (min(kSparseDc3, D - c0 - kSparseDc3) >> 2) : 0;
        const float* Ha = ring + bufa * g.unit_floats + 4 * half_lane;
        const float* Hb = ring + bufb * g.unit_floats + 4 * half_lane;
        float4 xa[kXSlots], xb[kXSlots];
#pragma unroll
        for (int sl = 0; sl < kXSlots; ++sl) {
            xa[sl] = xqa[sl];
            xb[sl] = xqb[sl];
            xqa[sl] = load_x(unit + 2, sl);                         // next pair's rows: in flight during this one
            xqb[sl] = load_x(unit + 3, sl);
        }
#ifdef DIGAT_TC_TIMING
        const long long tw0 = clock64();
#endif
        mbar_wait(&full[bufa], (uint32_t)(l / kSparseBufs) & 1u);
        if (pair) mbar_wait(&full[bufb], (uint32_t)((l + 1) / kSparseBufs) & 1u);
#ifdef DIGAT_TC_TIMING
        wait3 += clock64() - tw0;
#endif
        const bool on_a = half_lane < wqa, on_b = half_lane < wqb;
        constexpr int kMaxPasses = (kPairMaxNodes / 2 + kRowPairStep - 1) / kRowPairStep;
#pragma unroll
        for (int sl = 0; sl < kMaxPasses; ++sl) {                   // slots >= kXSlots load their residual directly
            const int rp = warp + sl * kRowPairStep;
            if (2 * rp >= n_act) break;
            const int i = row_of(sl);
            if (i >= 0 && on_a) {                                   // pruned rows are not in the list: neither read nor written
                const size_t xoff = ((size_t)src0 * n + i) * D + c0 + 4 * half_lane;
                const size_t yoff = ((size_t)b * n + i) * D + c0 + 4 * half_lane;
                float4 x0, x1 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (sl < kXSlots) {
                    x0 = xa[sl < kXSlots ? sl : 0];
                    x1 = xb[sl < kXSlots ? sl : 0];
                } else {
                    x0 = ldg_stream(reinterpret_cast<const float4*>(p.X + xoff));
                    if (on_b) x1 = ldg_stream(reinterpret_cast<const float4*>(p.X + xoff + kSparseDc3));
                }
                uint64_t o01 = 0ull, o23 = 0ull, p01 = 0ull, p23 = 0ull;
                const int e1 = rowptr[i + 1];
                int e = rowptr[i];
                for (; e + 2 <= e1; e += 2) {                                               // 2 neighbours x 2 units in flight
                    const float al0 = score[e], al1 = score[e + 1];                         // broadcast within the half warp
                    const uint32_t j0 = (meta[e] & 255u) * kSparseDc3, j1 = (meta[e + 1] & 255u) * kSparseDc3;
                    const float4 h0 = *reinterpret_cast<const float4*>(Ha + j0);
                    const float4 h1 = *reinterpret_cast<const float4*>(Ha + j1);
                    const float4 k0 = *reinterpret_cast<const float4*>(Hb + j0);            // (a stale / unused buffer when !on_b)
                    const float4 k1 = *reinterpret_cast<const float4*>(Hb + j1);
                    const uint64_t a0 = pack2(al0, al0), a1 = pack2(al1, al1);
                    o01 = fma2(a0, pack2(h0.x, h0.y), o01);
                    o23 = fma2(a0, pack2(h0.z, h0.w), o23);
                    p01 = fma2(a0, pack2(k0.x, k0.y), p01);
                    p23 = fma2(a0, pack2(k0.z, k0.w), p23);
                    o01 = fma2(a1, pack2(h1.x, h1.y), o01);
                    o23 = fma2(a1, pack2(h1.z, h1.w), o23);
                    p01 = fma2(a1, pack2(k1.x, k1.y), p01);
                    p23 = fma2(a1, pack2(k1.z, k1.w), p23);
                }
                if (e < e1) {
                    const float al = score[e];
                    const uint32_t j0 = (meta[e] & 255u) * kSparseDc3;
                    const float4 h = *reinterpret_cast<const float4*>(Ha + j0);
                    const float4 k = *reinterpret_cast<const float4*>(Hb + j0);
                    const uint64_t aa = pack2(al, al);
                    o01 = fma2(aa, pack2(h.x, h.y), o01);
                    o23 = fma2(aa, pack2(h.z, h.w), o23);
                    p01 = fma2(aa, pack2(k.x, k.y), p01);
                    p23 = fma2(aa, pack2(k.z, k.w), p23);
                }
                float4 y;
                unpack2(o01, y.x, y.y);
                unpack2(o23, y.z, y.w);
                if (kTrain && p.relu_mask_out != nullptr)                   // training: 1[alpha~ h > 0] for the backward
                    *reinterpret_cast<uchar4*>(p.relu_mask_out + yoff) = make_uchar4(y.x > 0.f, y.y > 0.f, y.z > 0.f, y.w > 0.f);
                y.x = fmaxf(y.x, 0.f) + x0.x;
                y.y = fmaxf(y.y, 0.f) + x0.y;
                y.z = fmaxf(y.z, 0.f) + x0.z;
                y.w = fmaxf(y.w, 0.f) + x0.w;
                stg_stream(reinterpret_cast<float4*>(p.Y + yoff), y);
                float* yc = p.Yc != nullptr ? p.Yc + (size_t)pos_s[i] * D + c0 + 4 * half_lane : nullptr;
                if (yc != nullptr) *reinterpret_cast<float4*>(yc) = y;      // compact copy for the next layer's projection
                if (on_b) {
                    unpack2(p01, y.x, y.y);
                    unpack2(p23, y.z, y.w);
                    if (kTrain && p.relu_mask_out != nullptr)
                        *reinterpret_cast<uchar4*>(p.relu_mask_out + yoff + kSparseDc3) = make_uchar4(y.x > 0.f, y.y > 0.f, y.z > 0.f, y.w > 0.f);
                    y.x = fmaxf(y.x, 0.f) + x1.x;
                    y.y = fmaxf(y.y, 0.f) + x1.y;
                    y.z = fmaxf(y.z, 0.f) + x1.z;
                    y.w = fmaxf(y.w, 0.f) + x1.w;
                    stg_stream(reinterpret_cast<float4*>(p.Y + yoff + kSparseDc3), y);
                    if (yc != nullptr) *reinterpret_cast<float4*>(yc + kSparseDc3) = y;
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&empty[bufa]);
            if (pair) mbar_arrive(&empty[bufb]);
        }
    }
#ifdef DIGAT_TC_TIMING
    if (tid == 0 && (b % 1000) == 7)
        printf("graph %d E=%d n_act=%d csr=%d: setup %lld (loads %lld), phase1 %lld (waiting on TMA %lld), softmax %lld, phase3 %lld (waiting %lld)\n",
               b, E, n_act, (int)have_csr, t1 - t0, ta - t0, t2 - t1, wait1, t3 - t2, clock64() - t3, wait3);
#endif
}

inline void sparse_geometry(int n, int D, int G, SparseGeom* g) {
    const int NB = G * n;
    g->G = G;
    g->nch1 = (D + kSparseDc1 - 1) / kSparseDc1;                 // (the launcher zeroes it for vanilla-GAT scores)
    g->nch3 = (D + kSparseDc3 - 1) / kSparseDc3;
    g->tile_floats = ((NB * kSparseDc1 * 4 + 1023) / 1024) * 1024 / 4;
    g->unit_floats = 2 * g->tile_floats;
    g->smem = (size_t)kSparseBufs * g->unit_floats * 4 + (size_t)2 * kSparseBufs * 8 + (size_t)2 * D * 4 +
              (size_t)NB * n * 4 + (size_t)(NB + 2) * 4 + (size_t)2 * NB * n + (size_t)((NB * n + 15) & ~15) + (size_t)2 * NB + (size_t)4 * NB + (size_t)NB + 8 + 64;
}

size_t graph_layer_fwd_sparse_smem(int n, int D) {
    SparseGeom g;
    sparse_geometry(n, D, 1, &g);
    return g.smem;
}

// Graphs per CTA for small graphs (not indexed): at most 128 nodes and the shared memory of DIGAT_SPARSE_MINCTAS CTAs per
// SM; among those the G that minimises (waves of CTAs) x (G + fixed per-CTA cost).
inline int sparse_graphs_per_cta(int n, int D, int B, int sm_count, size_t max_smem) {
    const size_t budget = (size_t)(227 * 1024) / DIGAT_SPARSE_MINCTAS - 1024 < max_smem ? (size_t)(227 * 1024) / DIGAT_SPARSE_MINCTAS - 1024 : max_smem;
    int best = 1;
    long best_cost = -1;
    for (int G = 1; G * n <= kPairMaxNodes && G <= B; ++G) {
        SparseGeom g;
        sparse_geometry(n, D, G, &g);
        if (G > 1 && g.smem > budget) break;
        const long ctas = (B + G - 1) / G, slots = (long)sm_count * DIGAT_SPARSE_MINCTAS;
        const long cost = ((ctas + slots - 1) / slots) * (G * (long)n + 24);     // ~24 nodes' worth of fixed work per CTA
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = G; }
    }
    return best;
}

int launch_graph_layer_fwd_sparse(const PairAttnArgs& args, int n_src, cudaStream_t st) {
    const DeviceInfo* di = device_info();
    if (!di) return fail(DIGAT_E_CUDA, "digat_graph_layer_fwd: no CUDA device");
    const bool indexed = args.px_index != nullptr;
    const bool training = args.alpha_out != nullptr || args.drop_keep != nullptr || args.relu_mask_out != nullptr;
    const int G = (indexed || args.adj_index != nullptr || training) ? 1
                  : sparse_graphs_per_cta(args.n, args.D, args.B, di->sm_count, (size_t)di->max_smem_optin);
    SparseGeom g;
    sparse_geometry(args.n, args.D, G, &g);
    g.stagger_lo = di->sm_count;
    g.stagger_hi = 2 * di->sm_count;
    DIGAT_REQUIRE(g.smem <= (size_t)di->max_smem_optin, "digat_graph_layer_fwd(sparse): needs %zu B shared memory", g.smem);
    CUtensorMap map1, map3;
    int rc;
    const int64_t src_graphs = indexed ? n_src : args.B;
    const int p_cols = args.gat_s != nullptr ? args.D : 3 * args.D;     // vanilla GAT: P = h only, no U | K2 loads
    if (args.gat_s != nullptr) g.nch1 = 0;
    if ((rc = make_tensor_map_2d(&map1, args.P, src_graphs * args.n, p_cols, args.ldp, G * args.n, kSparseDc1, CU_TENSOR_MAP_SWIZZLE_128B)) != DIGAT_OK) return rc;
    if ((rc = make_tensor_map_2d(&map3, args.P, src_graphs * args.n, p_cols, args.ldp, G * args.n, kSparseDc3, CU_TENSOR_MAP_SWIZZLE_NONE)) != DIGAT_OK) return rc;
    const int grid = (args.B + G - 1) / G;
    if (indexed) {
        if (int rc_ = ensure_dynamic_smem(graph_layer_fwd_sparse_kernel<true, false>, (size_t)(g.smem))) return rc_;
        graph_layer_fwd_sparse_kernel<true, false><<<grid, kSparseThreads, g.smem, st>>>(map1, map3, args, g);
    } else if (training) {
        if (int rc_ = ensure_dynamic_smem(graph_layer_fwd_sparse_kernel<false, false, true>, (size_t)(g.smem))) return rc_;
        graph_layer_fwd_sparse_kernel<false, false, true><<<grid, kSparseThreads, g.smem, st>>>(map1, map3, args, g);
    } else if (G == 1) {
        if (int rc_ = ensure_dynamic_smem(graph_layer_fwd_sparse_kernel<false, false>, (size_t)(g.smem))) return rc_;
        graph_layer_fwd_sparse_kernel<false, false><<<grid, kSparseThreads, g.smem, st>>>(map1, map3, args, g);
    } else {
        if (int rc_ = ensure_dynamic_smem(graph_layer_fwd_sparse_kernel<false, true>, (size_t)(g.smem))) return rc_;
        graph_layer_fwd_sparse_kernel<false, true><<<grid, kSparseThreads, g.smem, st>>>(map1, map3, args, g);
    }
    return check_launch("digat_graph_layer_fwd(sparse)");
}

// Vanilla-GAT layer of the ablation encoders (reference graphEncoders.py:494-503 / 511-520 and the two mixed variants):
//   e_ij = leaky_relu(a1 . h_j + a2 . h_i);  alpha = softmax_j(mask(e));  Y = relu(alpha h) + X
// on the edge-driven kernel: the score of an edge is one add of two precomputed dot products (gat_s), so the kernel
// streams h only.
inline int launch_gat_layer_fwd(const float* Hm, int ldh, const float* s12, const uint8_t* adj, const float* X, float* Y,
                                int B, int n, int D, cudaStream_t st) {
    if (B == 0) return DIGAT_OK;
    DIGAT_REQUIRE(Hm && s12 && adj && X && Y, "digat_gat_layer_fwd: null pointer");
    DIGAT_REQUIRE(B >= 0 && n >= 1 && n <= kPairMaxNodes, "digat_gat_layer_fwd: n=%d outside [1,%d]", n, kPairMaxNodes);
    DIGAT_REQUIRE(D >= 4 && (D & 3) == 0 && D <= 1024, "digat_gat_layer_fwd: D=%d must be a multiple of 4 in [4,1024]", D);
    DIGAT_REQUIRE((ldh & 3) == 0 && ldh >= D, "digat_gat_layer_fwd: ldh=%d must be a multiple of 4 and >= D", ldh);
    DIGAT_REQUIRE(aligned16(Hm) && aligned16(X) && aligned16(Y), "digat_gat_layer_fwd: pointers must be 16-byte aligned");
    const DeviceInfo* di = device_info();
    if (!di) return fail(DIGAT_E_CUDA, "digat_gat_layer_fwd: no CUDA device");
    DIGAT_REQUIRE(graph_layer_fwd_sparse_smem(n, D) <= (size_t)di->max_smem_optin,
                  "digat_gat_layer_fwd: a graph of %d nodes x %d features does not fit one CTA", n, D);
    PairAttnArgs args{Hm, ldh, nullptr, adj, X, Y, B, n, D, nullptr, 1.f, nullptr, nullptr, nullptr,
                      nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr};
    args.gat_s = s12;
    return launch_graph_layer_fwd_sparse(args, B, st);
}

// Training forward of the vanilla-GAT layer: the same launch with the attention dropout and the per-edge score / alpha / relu
// mask outputs of the edge-driven training path (precomputed CSR required), for digat_gat_layer_bwd_csr.
inline int launch_gat_layer_train_fwd(const float* Hm, int ldh, const float* s12, const uint8_t* adj, const float* X, float* Y,
                                      int B, int n, int D, const uint8_t* drop_keep, float drop_scale, float* edge_score,
                                      float* edge_alpha, uint8_t* relu_mask, const uint16_t* csr_rowptr, const uint16_t* csr_meta,
                                      cudaStream_t st) {
    if (B == 0) return DIGAT_OK;
    DIGAT_REQUIRE(Hm && s12 && adj && X && Y && edge_score && edge_alpha && relu_mask && csr_rowptr && csr_meta,
                  "digat_gat_layer_train_fwd: null pointer");
    DIGAT_REQUIRE(B >= 0 && n >= 1 && n <= kPairMaxNodes, "digat_gat_layer_train_fwd: n=%d outside [1,%d]", n, kPairMaxNodes);
    DIGAT_REQUIRE(D >= 4 && (D & 3) == 0 && D <= 1024, "digat_gat_layer_train_fwd: D=%d must be a multiple of 4 in [4,1024]", D);
    DIGAT_REQUIRE((ldh & 3) == 0 && ldh >= D, "digat_gat_layer_train_fwd: ldh=%d must be a multiple of 4 and >= D", ldh);
    DIGAT_REQUIRE(aligned16(Hm) && aligned16(X) && aligned16(Y), "digat_gat_layer_train_fwd: pointers must be 16-byte aligned");
    const DeviceInfo* di = device_info();
    if (!di) return fail(DIGAT_E_CUDA, "digat_gat_layer_train_fwd: no CUDA device");
    DIGAT_REQUIRE(graph_layer_fwd_sparse_smem(n, D) <= (size_t)di->max_smem_optin,
                  "digat_gat_layer_train_fwd: a graph of %d nodes x %d features does not fit one CTA", n, D);
    PairAttnArgs args{Hm, ldh, nullptr, adj, X, Y, B, n, D, drop_keep, drop_scale, edge_score, edge_alpha, relu_mask,
                      nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr};
    args.gat_s = s12;
    args.csr_rowptr = csr_rowptr;
    args.csr_meta = csr_meta;
    return launch_graph_layer_fwd_sparse(args, B, st);
}

}  // namespace digat
