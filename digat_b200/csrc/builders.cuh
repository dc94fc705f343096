// Integer / byte builders either side of the encoder: all results bit-exact.
//   user_graph_kernel       reference MIND_corpus.py:143-176  (user-history graph, category mask, segment ids)
//   sag_bfs_kernel          reference construct_SAG.py:449-485 (semantic-augmented graph: BFS over similar news)
//   rank_impressions_kernel reference util.py:70-80           (per-impression rank of every pair, stable descending)
//   impression_metrics_kernel reference evaluate.py:32-89     (AUC / MRR / nDCG@5 / nDCG@10 of score = 1/rank)
// These are HBM-bound byte kernels (the user-graph builder writes 4.6 KB per behaviour from 200 B of input);
// nothing here is GEMM-shaped.
#pragma once
#include "common.cuh"

namespace digat {

constexpr int kBuilderThreads = 256;
constexpr int kBuilderMaxH = 512;      // history slots kept in shared memory
constexpr int kBuilderMaxC = 256;      // categories

// One CTA per behaviour.  Element (i, j) of the [n_u, n_u] graph, n_u = H + C (nodes 0..H-1 clicked news, H.. topics):
//   i == j                      -> 1                                     (np.identity, MIND_corpus.py:145)
//   both news                   -> both valid and same category         (:168-170)
//   news i, topic c             -> i valid and category(i) == c         (:163-164)
//   topic c, topic c'           -> both categories present              (:171-173; distinct categories only: c != c')
// kWord: n_u^2 is a multiple of 4, so a thread produces 4 consecutive bytes with one 32-bit store.
template <bool kWord>
__global__ void __launch_bounds__(kBuilderThreads)
user_graph_kernel(const int32_t* __restrict__ hist_cat, const int32_t* __restrict__ hist_len, uint8_t* __restrict__ graph,
                  uint8_t* __restrict__ cmask, int64_t* __restrict__ cidx, int H, int C, int32_t* __restrict__ err_flag) {
    __shared__ int cat_s[kBuilderMaxH];          // category of slot t, C for padding
    __shared__ int present_s[kBuilderMaxC + 1];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n_u = H + C;
    int len = hist_len[b];
    if (len < 0 || len > H) {
        if (tid == 0 && err_flag) atomicOr(err_flag, 1);
        len = min(max(len, 0), H);
    }
    for (int c = tid; c <= C; c += kBuilderThreads) present_s[c] = 0;
    __syncthreads();
    for (int t = tid; t < H; t += kBuilderThreads) {
        int c = C;
        if (t < len) {
            c = hist_cat[(size_t)b * H + t];
            if (c < 0 || c >= C) {                 // the reference would raise IndexError on cmask[c]
                if (err_flag) atomicOr(err_flag, 1);
                c = C;
            } else {
                present_s[c] = 1;                  // benign race: every writer stores 1
            }
        }
        cat_s[t] = c;
        cidx[(size_t)b * H + t] = c;
    }
    __syncthreads();
    for (int c = tid; c <= C; c += kBuilderThreads) cmask[(size_t)b * (C + 1) + c] = (c < C && present_s[c]) ? 1 : 0;

    auto element = [&](int i, int j) -> uint32_t {
        if (i == j) return 1u;
        if (i < H && j < H) return (cat_s[i] == cat_s[j] && cat_s[i] < C) ? 1u : 0u;
        if (i < H) return cat_s[i] == j - H ? 1u : 0u;            // padding slots hold C, never equal to a topic id < C
        if (j < H) return cat_s[j] == i - H ? 1u : 0u;
        return (present_s[i - H] && present_s[j - H]) ? 1u : 0u;
    };
    const int n2 = n_u * n_u;
    uint8_t* g = graph + (size_t)b * n2;
    if (kWord) {
        uint32_t* g4 = reinterpret_cast<uint32_t*>(g);
        for (int w = tid; w < n2 / 4; w += kBuilderThreads) {
            int i = (4 * w) / n_u, j = 4 * w - i * n_u;
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v |= element(i, j) << (8 * k);
                if (++j == n_u) { j = 0; ++i; }
            }
            g4[w] = v;
        }
    } else {
        for (int e = tid; e < n2; e += kBuilderThreads) {
            const int i = e / n_u;
            g[e] = (uint8_t)element(i, e - i * n_u);
        }
    }
}

inline int launch_build_user_graphs(const int32_t* hist_cat, const int32_t* hist_len, uint8_t* graph, uint8_t* cmask,
                                    int64_t* cidx, int64_t N, int H, int C, int32_t* err_flag, cudaStream_t st) {
    if (N <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(hist_cat && hist_len && graph && cmask && cidx, "digat_build_user_graphs: null pointer");
    DIGAT_REQUIRE(H >= 1 && H <= kBuilderMaxH && C >= 1 && C <= kBuilderMaxC,
                  "digat_build_user_graphs: H=%d must be in [1,%d], C=%d in [1,%d]", H, kBuilderMaxH, C, kBuilderMaxC);
    DIGAT_REQUIRE(N <= 0x7fffffff, "digat_build_user_graphs: too many behaviours for one launch");
    const int n2 = (H + C) * (H + C);
    if ((n2 & 3) == 0 && (reinterpret_cast<uintptr_t>(graph) & 3) == 0)
        user_graph_kernel<true><<<(unsigned)N, kBuilderThreads, 0, st>>>(hist_cat, hist_len, graph, cmask, cidx, H, C, err_flag);
    else
        user_graph_kernel<false><<<(unsigned)N, kBuilderThreads, 0, st>>>(hist_cat, hist_len, graph, cmask, cidx, H, C, err_flag);
    return check_launch("digat_build_user_graphs");
}

// Node pruning for the user graph (inference): active[g, i] = 0 iff nothing can observe the layer output of node i:
//   * no OTHER node attends to it (column i of the adjacency is empty off the diagonal), and
//   * no context reads it: i >= H (topic nodes are never pooled, graphEncoders.py:125), or its history slot belongs to a
//     bucket c = cidx[g,i] whose topic embedding is masked out of the user-level attention (cmask[g,c] == 0; in MIND data
//     the padding bucket S-1) -- unless EVERY bucket is masked: the -1e9 fill then gives a uniform softmax over all S
//     buckets (layers.py:202), which reads them all (users with an empty history).
// A graph with an edge-less row keeps every node (that row's softmax is uniform over ALL nodes).
// For MIND-shaped graphs the inactive nodes are the padded history slots and the categories the user never clicked.
__global__ void __launch_bounds__(128)
user_active_rows_kernel(const uint8_t* __restrict__ adj, const int32_t* __restrict__ adj_index,
                        const int64_t* __restrict__ cidx, const uint8_t* __restrict__ cmask, uint8_t* __restrict__ active,
                        uint8_t* __restrict__ pooled_out, int n, int H, int S) {
    __shared__ int col_used[128];
    __shared__ int any_empty, any_bucket;
    const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t src = adj_index != nullptr ? (size_t)adj_index[g] : (size_t)g;
    const uint8_t* a = adj + src * n * n;
    if (tid < n) col_used[tid] = 0;
    if (tid == 0) { any_empty = 0; any_bucket = 0; }
    __syncthreads();
    for (int c = tid; c < S; c += 128)
        if (cmask[(size_t)g * S + c] != 0) any_bucket = 1;
    for (int i = warp; i < n; i += 4) {
        bool row_any = false;
        for (int j = lane; j < n; j += 32) {
            const bool on = a[i * n + j] != 0;
            row_any |= on;
            if (on && j != i) col_used[j] = 1;          // benign race: every writer stores 1
        }
        if (!__any_sync(0xffffffffu, row_any) && lane == 0) any_empty = 1;
    }
    __syncthreads();
    if (tid < n) {
        bool pooled = false;                                // does a context read this node's output?
        if (tid < H) {
            const int64_t c = cidx[src * H + tid];
            pooled = !any_bucket || c < 0 || c >= S || cmask[(size_t)g * S + c] != 0;   // bad ids: leave to the segment kernel
        }
        active[(size_t)g * n + tid] = (any_empty || col_used[tid] || pooled) ? 1 : 0;
        if (pooled_out != nullptr) pooled_out[(size_t)g * n + tid] = pooled ? 1 : 0;   // rows a context reads directly
    }
}

// The same for news graphs (SAG): node i is observable iff another node attends to it, or the news context reads it:
// i == 0 (the "local" context, graphEncoders.py:110), mask[g,i] != 0 (pooled by the candidate attention) or every mask
// entry is 0 (uniform softmax over all nodes).  In SAG data the inactive nodes are the unused BFS slots.
__global__ void __launch_bounds__(128)
news_active_rows_kernel(const uint8_t* __restrict__ adj, const uint8_t* __restrict__ mask, uint8_t* __restrict__ active, int n) {
    __shared__ int col_used[128];
    __shared__ int any_empty, any_mask;
    const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t* a = adj + (size_t)g * n * n;
    if (tid < n) col_used[tid] = 0;
    if (tid == 0) { any_empty = 0; any_mask = 0; }
    __syncthreads();
    if (tid < n && mask[(size_t)g * n + tid] != 0) any_mask = 1;
    for (int i = warp; i < n; i += 4) {
        bool row_any = false;
        for (int j = lane; j < n; j += 32) {
            const bool on = a[i * n + j] != 0;
            row_any |= on;
            if (on && j != i) col_used[j] = 1;
        }
        if (!__any_sync(0xffffffffu, row_any) && lane == 0) any_empty = 1;
    }
    __syncthreads();
    if (tid < n)
        active[(size_t)g * n + tid] = (any_empty || col_used[tid] || tid == 0 || !any_mask || mask[(size_t)g * n + tid] != 0) ? 1 : 0;
}

inline int launch_news_active_rows(const uint8_t* adj, const uint8_t* mask, uint8_t* active, int64_t G, int n, cudaStream_t st) {
    if (G <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(adj && mask && active, "digat_news_active_rows: null pointer");
    DIGAT_REQUIRE(n >= 1 && n <= 128, "digat_news_active_rows: bad n");
    DIGAT_REQUIRE(G <= 0x7fffffff, "digat_news_active_rows: too many graphs for one launch");
    news_active_rows_kernel<<<(unsigned)G, 128, 0, st>>>(adj, mask, active, n);
    return check_launch("digat_news_active_rows");
}

inline int launch_user_active_rows(const uint8_t* adj, const int32_t* adj_index, const int64_t* cidx, const uint8_t* cmask,
                                   uint8_t* active, uint8_t* pooled_out, int64_t G, int n, int H, int S, cudaStream_t st) {
    if (G <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(adj && cidx && cmask && active, "digat_user_active_rows: null pointer");
    DIGAT_REQUIRE(n >= 1 && n <= 128 && H >= 0 && H <= n && S >= 1, "digat_user_active_rows: bad n / H / S");
    DIGAT_REQUIRE(G <= 0x7fffffff, "digat_user_active_rows: too many graphs for one launch");
    user_active_rows_kernel<<<(unsigned)G, 128, 0, st>>>(adj, adj_index, cidx, cmask, active, pooled_out, n, H, S);
    return check_launch("digat_user_active_rows");
}

// CSR of a graph's adjacency, restricted to the evaluated rows (row_active), in the form the edge-driven layer kernel uses:
// rowptr[i+1] = end of row i's edge range (bit 15: the row has no edge at all -> uniform softmax over every node, the
// reference's all -1e9 row), meta[e] = neighbour | row << 8.  Built ONCE per batch and graph; the layer kernel used to rebuild
// it from the adjacency bytes in every layer and for every pair sharing the graph (12 % of its instructions, ~15k of ~80k
// cycles per graph).  One CTA of four warps per graph, a warp per row (ballot + popc), one warp scan for the row pointers.
__global__ void __launch_bounds__(128)
graph_csr_kernel(const uint8_t* __restrict__ adj, const int32_t* __restrict__ adj_index, const uint8_t* __restrict__ row_active,
                 uint16_t* __restrict__ rowptr_out, uint16_t* __restrict__ meta_out, uint16_t* __restrict__ colptr_out,
                 uint16_t* __restrict__ cedge_out, int n) {
    extern __shared__ uint16_t eid[];            // [n*n] edge id of (i,j) or 0xFFFF (only when the transpose is requested)
    __shared__ int rp[130];
    __shared__ int cp[130];
    __shared__ uint8_t uni[128];
    const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t* a = adj + (adj_index != nullptr ? (size_t)adj_index[g] : (size_t)g) * n * n;
    const uint8_t* act = row_active != nullptr ? row_active + (size_t)g * n : nullptr;
    const bool want_t = colptr_out != nullptr;
    const int n32 = (n + 31) & ~31;
    for (int i = warp; i < n; i += 4) {
        const bool dead = act != nullptr && act[i] == 0;
        int deg = 0;
        if (!dead)
            for (int j = lane; j < n32; j += 32)
                deg += __popc(__ballot_sync(0xffffffffu, j < n && a[i * n + j] != 0));
        if (lane == 0) {
            uni[i] = deg == 0 && !dead;
            rp[i + 1] = dead ? 0 : (deg == 0 ? n : deg);
        }
    }
    if (want_t)
        for (int e = tid; e < n * n; e += 128) eid[e] = 0xFFFFu;
    __syncthreads();
    if (warp == 0) {
        int run = 0;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            int v = i < n ? rp[i + 1] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            if (i < n) rp[i + 1] = run + v;
            run += __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) rp[0] = 0;
    }
    __syncthreads();
    uint16_t* rpo = rowptr_out + (size_t)g * (n + 1);
    uint16_t* mo = meta_out + (size_t)g * n * n;
    for (int i = tid; i <= n; i += 128) rpo[i] = (uint16_t)(rp[i] | ((i > 0 && uni[i - 1]) ? 0x8000 : 0));
    for (int i = warp; i < n; i += 4) {
        if (rp[i + 1] == rp[i]) continue;
        const int e0 = rp[i];
        const bool u = uni[i] != 0;
        int filled = 0;
        for (int j = lane; j < n32; j += 32) {
            const bool on = j < n && (u || a[i * n + j] != 0);
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (on) {
                const int e = e0 + filled + __popc(m & ((1u << lane) - 1u));
                mo[e] = (uint16_t)(j | (i << 8));
                if (want_t) eid[i * n + j] = (uint16_t)e;
            }
            filled += __popc(m);
        }
    }
    if (!want_t) return;
    // transpose (CSC): the edges of column j in ascending row order, as ids into the CSR arrays (the backward pass sums
    // over the incoming edges of a node: dh_j, dU_j)
    __syncthreads();
    for (int j = warp; j < n; j += 4) {
        int cnt = 0;
        for (int i = lane; i < n32; i += 32)
            cnt += __popc(__ballot_sync(0xffffffffu, i < n && eid[i * n + j] != 0xFFFFu));
        if (lane == 0) cp[j + 1] = cnt;
    }
    __syncthreads();
    if (warp == 0) {
        int run = 0;
        for (int base = 0; base < n; base += 32) {
            const int j = base + lane;
            int v = j < n ? cp[j + 1] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            if (j < n) cp[j + 1] = run + v;
            run += __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) cp[0] = 0;
    }
    __syncthreads();
    uint16_t* cpo = colptr_out + (size_t)g * (n + 1);
    uint16_t* ceo = cedge_out + (size_t)g * n * n;
    for (int j = tid; j <= n; j += 128) cpo[j] = (uint16_t)cp[j];
    for (int j = warp; j < n; j += 4) {
        int filled = 0;
        for (int i = lane; i < n32; i += 32) {
            const uint16_t e = i < n ? eid[i * n + j] : (uint16_t)0xFFFFu;
            const bool on = e != 0xFFFFu;
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (on) ceo[cp[j] + filled + __popc(m & ((1u << lane) - 1u))] = e;
            filled += __popc(m);
        }
    }
}

inline int launch_build_graph_csr(const uint8_t* adj, const int32_t* adj_index, const uint8_t* row_active, uint16_t* rowptr,
                                  uint16_t* meta, uint16_t* colptr, uint16_t* cedge, int64_t G, int n, cudaStream_t st) {
    if (G <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(adj && rowptr && meta, "digat_build_graph_csr: null pointer");
    DIGAT_REQUIRE((colptr == nullptr) == (cedge == nullptr), "digat_build_graph_csr: colptr and cedge go together");
    DIGAT_REQUIRE(n >= 1 && n <= 128 && G < (1LL << 31), "digat_build_graph_csr: n=%d outside [1,128]", n);
    const size_t smem = colptr != nullptr ? (size_t)n * n * sizeof(uint16_t) : 0;
    graph_csr_kernel<<<(unsigned)G, 128, smem, st>>>(adj, adj_index, row_active, rowptr, meta, colptr, cedge, n);
    return check_launch("digat_build_graph_csr");
}

// Stream compaction of up to four flag lists laid out back to back in `flags` (their inclusive prefix sums in `csum`):
// for list k covering flat positions [lo[k], lo[k] + size[k]) with base[k] set flags before it,
//   pos_k[r] = csum[lo + r] - 1 - base        (rank of position r among the set flags of ITS list; valid where set)
//   ids_k[pos_k[r]] = r                        for every set flag
// One launch builds all lists; the counts are known to the host (it sized ids_k), so nothing here synchronises.
struct CompactLists {
    int n_lists;
    int64_t lo[4], size[4];
    int32_t base[4];
    int32_t* ids[4];
    int32_t* pos[4];      // may be null (list whose ranks nobody needs)
};

__global__ void compact_lists_kernel(const uint8_t* __restrict__ flags, const int32_t* __restrict__ csum, CompactLists c) {
    const int64_t total = c.lo[c.n_lists - 1] + c.size[c.n_lists - 1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int k = 0;
#pragma unroll
        for (int t = 1; t < 4; ++t)
            if (t < c.n_lists && i >= c.lo[t]) k = t;
        const int64_t r = i - c.lo[k];
        const int32_t p = csum[i] - 1 - c.base[k];
        if (c.pos[k] != nullptr) c.pos[k][r] = p;
        if (flags[i] != 0) c.ids[k][p] = (int32_t)r;
    }
}

inline int launch_compact_lists(const uint8_t* flags, const int32_t* csum, int n_lists, const int64_t* lo, const int64_t* size,
                                const int32_t* base, int32_t* const* ids, int32_t* const* pos, cudaStream_t st) {
    DIGAT_REQUIRE(flags && csum && lo && size && base && ids && pos, "digat_compact_lists: null pointer");
    DIGAT_REQUIRE(n_lists >= 1 && n_lists <= 4, "digat_compact_lists: 1..4 lists");
    CompactLists c;
    c.n_lists = n_lists;
    for (int k = 0; k < 4; ++k) {
        const bool on = k < n_lists;
        c.lo[k] = on ? lo[k] : 0; c.size[k] = on ? size[k] : 0; c.base[k] = on ? base[k] : 0;
        c.ids[k] = on ? ids[k] : nullptr; c.pos[k] = on ? pos[k] : nullptr;
        DIGAT_REQUIRE(!on || (size[k] >= 0 && (k == 0 ? lo[k] == 0 : lo[k] == lo[k - 1] + size[k - 1])),
                      "digat_compact_lists: lists must be laid out back to back");
        DIGAT_REQUIRE(!on || ids[k] != nullptr || size[k] == 0, "digat_compact_lists: null ids");
    }
    const int64_t total = c.lo[n_lists - 1] + c.size[n_lists - 1];
    if (total <= 0) return DIGAT_OK;
    const DeviceInfo* di = device_info();
    if (!di) return fail(DIGAT_E_CUDA, "digat_compact_lists: no CUDA device");
    const int64_t blocks = (total + 255) / 256;
    compact_lists_kernel<<<(unsigned)(blocks < 8 * di->sm_count ? blocks : 8 * di->sm_count), 256, 0, st>>>(flags, csum, c);
    return check_launch("digat_compact_lists");
}

// ---------------------------------------------------------------------------------------------------- SAG BFS
// One warp per news.  The queue (node ids, depths) and the n x n adjacency live in shared memory; the similar-news
// lists come as CSR (offsets, neighbour index, cosine as double: the reference compares python floats).
// The BFS itself is sequential (construct_SAG.py:459-484): lane 0 walks the neighbour list, all lanes search the
// queue for "already present" with ballots.
constexpr int kSagWarps = 4;
constexpr int kSagMaxNodes = 128;

__global__ void __launch_bounds__(kSagWarps * 32)
sag_bfs_kernel(const int64_t* __restrict__ sim_off, const int32_t* __restrict__ sim_idx, const double* __restrict__ sim_cos,
               int32_t* __restrict__ node_id, uint8_t* __restrict__ graph, uint8_t* __restrict__ mask, int n_news,
               int top_M, int hop, int n, double threshold, int32_t* __restrict__ err_flag) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n2 = n * n;
    const int g_bytes = (n2 + 15) & ~15;
    uint8_t* g_s = smem_raw + (size_t)warp * g_bytes;
    int* node_s = reinterpret_cast<int*>(smem_raw + (size_t)kSagWarps * g_bytes) + warp * 2 * kSagMaxNodes;
    int* depth_s = node_s + kSagMaxNodes;
    const int news = blockIdx.x * kSagWarps + warp;
    if (news >= n_news) return;

    for (int e = lane; e < g_bytes / 4; e += 32) reinterpret_cast<uint32_t*>(g_s)[e] = 0u;
    for (int e = lane; e < n; e += 32) { node_s[e] = 0; depth_s[e] = 0; }
    __syncwarp();
    int rear = 0;
    if (news >= 1) {                                   // row 0 is the padding news: all zero except mask[0,0]
        if (lane == 0) node_s[0] = news;
        __syncwarp();
        rear = 1;
        for (int head = 0; head < rear; ++head) {
            const int d = depth_s[head];
            if (d == hop) continue;
            const int cur = node_s[head];
            const int64_t lo = sim_off[cur], hi = sim_off[cur + 1];
            for (int64_t k = lo; k < hi; ++k) {
                const double cs = sim_cos[k];
                if (d > 0 && (cs < threshold || (int)(k - lo) == top_M - 1)) break;
                const int other = sim_idx[k];
                if (other < 0 || other >= n_news) {
                    if (lane == 0 && err_flag) atomicOr(err_flag, 1);
                    continue;
                }
                int pos = -1;
                for (int base = 0; base < rear; base += 32) {
                    const unsigned hit = __ballot_sync(0xffffffffu, base + lane < rear && node_s[base + lane] == other);
                    if (hit) { pos = base + __ffs(hit) - 1; break; }
                }
                if (pos < 0) {
                    if (rear >= n) {                   // the reference would raise IndexError
                        if (lane == 0 && err_flag) atomicOr(err_flag, 2);
                        continue;
                    }
                    pos = rear;
                    if (lane == 0) { node_s[rear] = other; depth_s[rear] = d + 1; }
                    ++rear;
                }
                if (lane == 0) { g_s[head * n + pos] = 1; g_s[pos * n + head] = 1; }
                __syncwarp();
            }
        }
    }
    __syncwarp();
    // coalesced write-out
    for (int e = lane; e < n; e += 32) {
        node_id[(size_t)news * n + e] = node_s[e];
        mask[(size_t)news * n + e] = (e == 0 || e < rear) ? 1 : 0;
    }
    uint8_t* g = graph + (size_t)news * n2;
    for (int e = lane; e < n2; e += 32) g[e] = g_s[e];
}

inline int launch_sag_bfs(const int64_t* sim_off, const int32_t* sim_idx, const double* sim_cos, int32_t* node_id,
                          uint8_t* graph, uint8_t* mask, int n_news, int top_M, int hop, int n_nodes, double threshold,
                          int32_t* err_flag, cudaStream_t st) {
    if (n_news <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(sim_off && node_id && graph && mask, "digat_sag_bfs: null pointer");
    DIGAT_REQUIRE(n_nodes >= 1 && n_nodes <= kSagMaxNodes, "digat_sag_bfs: n_nodes=%d outside [1,%d]", n_nodes, kSagMaxNodes);
    DIGAT_REQUIRE(top_M >= 1 && hop >= 0, "digat_sag_bfs: bad top_M / hop");
    const size_t smem = (size_t)kSagWarps * (((size_t)n_nodes * n_nodes + 15) & ~(size_t)15) +
                        (size_t)kSagWarps * 2 * kSagMaxNodes * sizeof(int);
    if (int rc_ = ensure_dynamic_smem(sag_bfs_kernel, (size_t)(smem))) return rc_;
    sag_bfs_kernel<<<(n_news + kSagWarps - 1) / kSagWarps, kSagWarps * 32, smem, st>>>(
        sim_off, sim_idx, sim_cos, node_id, graph, mask, n_news, top_M, hop, n_nodes, threshold, err_flag);
    return check_launch("digat_sag_bfs");
}

// ---------------------------------------------------------------------------------------------------- ranking
// rank_i = 1 + #{j : s_j > s_i} + #{j < i : s_j == s_i}: the position of pair i in a STABLE descending sort of its
// impression (util.py:72-76 sorts [score, index] rows with a stable list sort keyed on -score).  One warp per impression.
constexpr int kRankWarps = 8;

__global__ void __launch_bounds__(kRankWarps * 32)
rank_impressions_kernel(const float* __restrict__ scores, const int64_t* __restrict__ offsets, int32_t* __restrict__ ranks,
                        int64_t n_imp) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t imp = (int64_t)blockIdx.x * kRankWarps + warp;
    if (imp >= n_imp) return;
    const int64_t lo = offsets[imp];
    const int m = (int)(offsets[imp + 1] - lo);
    const float* s = scores + lo;
    for (int base = 0; base < m; base += 32) {
        const int i = base + lane;
        const float si = i < m ? s[i] : 0.f;
        int r = 1;
        for (int j0 = 0; j0 < m; j0 += 32) {
            const float mine = j0 + lane < m ? s[j0 + lane] : 0.f;
            const int cnt = min(32, m - j0);
            for (int k = 0; k < cnt; ++k) {
                const float sj = __shfl_sync(0xffffffffu, mine, k);
                const int j = j0 + k;
                r += (sj > si || (sj == si && j < i)) ? 1 : 0;
            }
        }
        if (i < m) ranks[lo + i] = r;
    }
}

inline int launch_rank_impressions(const float* scores, const int64_t* offsets, int32_t* ranks, int64_t n_imp,
                                   cudaStream_t st) {
    if (n_imp <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(scores && offsets && ranks, "digat_rank_impressions: null pointer");
    rank_impressions_kernel<<<(unsigned)((n_imp + kRankWarps - 1) / kRankWarps), kRankWarps * 32, 0, st>>>(
        scores, offsets, ranks, n_imp);
    return check_launch("digat_rank_impressions");
}

// Per-impression metrics of y_score = 1/rank (evaluate.py:73-80).  Ranks are a permutation of 1..m, so there are no
// ties: AUC = #{(pos, neg) : rank_pos < rank_neg} / (n_pos n_neg) (integer count, one division),
// MRR = sum_pos 1/rank / n_pos, nDCG@k = sum_{pos, rank <= k} 1/log2(rank + 1) / sum_{t < min(k, n_pos)} 1/log2(t + 2).
// out[imp] = {auc, mrr, ndcg5, ndcg10} in double; valid[imp] = 0 for an empty impression (skipped by the reference,
// evaluate.py:70) and 2 when only one class is present (roc_auc_score raises) -- the caller turns 2 into an error.
__global__ void __launch_bounds__(kRankWarps * 32)
impression_metrics_kernel(const int32_t* __restrict__ ranks, const uint8_t* __restrict__ labels,
                          const int64_t* __restrict__ offsets, double* __restrict__ out, uint8_t* __restrict__ valid,
                          int64_t n_imp) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t imp = (int64_t)blockIdx.x * kRankWarps + warp;
    if (imp >= n_imp) return;
    const int64_t lo = offsets[imp];
    const int m = (int)(offsets[imp + 1] - lo);
    long long n_pos = 0, below = 0;                   // below = sum over positives of (m - rank)
    double rr = 0.0, d5 = 0.0, d10 = 0.0;
    for (int i = lane; i < m; i += 32) {
        if (labels[lo + i] != 0) {
            const int r = ranks[lo + i];
            ++n_pos;
            below += m - r;
            rr += 1.0 / (double)r;
            const double g = 1.0 / log2((double)r + 1.0);
            if (r <= 5) d5 += g;
            if (r <= 10) d10 += g;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_pos += __shfl_xor_sync(0xffffffffu, n_pos, o);
        below += __shfl_xor_sync(0xffffffffu, below, o);
        rr += __shfl_xor_sync(0xffffffffu, rr, o);
        d5 += __shfl_xor_sync(0xffffffffu, d5, o);
        d10 += __shfl_xor_sync(0xffffffffu, d10, o);
    }
    if (lane != 0) return;
    double* o4 = out + 4 * imp;
    const long long n_neg = m - n_pos;
    if (m == 0) { valid[imp] = 0; o4[0] = o4[1] = o4[2] = o4[3] = 0.0; return; }
    if (n_pos == 0 || n_neg == 0) { valid[imp] = 2; o4[0] = o4[1] = o4[2] = o4[3] = 0.0; return; }
    double i5 = 0.0, i10 = 0.0;
    for (int t = 0; t < 10 && t < n_pos; ++t) {
        const double g = 1.0 / log2((double)t + 2.0);
        if (t < 5) i5 += g;
        i10 += g;
    }
    const long long wins = below - n_pos * (n_pos - 1) / 2;      // negatives ranked below each positive, summed
    o4[0] = (double)wins / ((double)n_pos * (double)n_neg);
    o4[1] = rr / (double)n_pos;
    o4[2] = d5 / i5;
    o4[3] = d10 / i10;
    valid[imp] = 1;
}

inline int launch_impression_metrics(const int32_t* ranks, const uint8_t* labels, const int64_t* offsets, double* out,
                                     uint8_t* valid, int64_t n_imp, cudaStream_t st) {
    if (n_imp <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(ranks && labels && offsets && out && valid, "digat_impression_metrics: null pointer");
    impression_metrics_kernel<<<(unsigned)((n_imp + kRankWarps - 1) / kRankWarps), kRankWarps * 32, 0, st>>>(
        ranks, labels, offsets, out, valid, n_imp);
    return check_launch("digat_impression_metrics");
}

}  // namespace digat
