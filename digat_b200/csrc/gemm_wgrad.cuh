// Weight-gradient GEMM and the small reductions of the backward pass (exact fp32, CUDA cores, deterministic):
//   dW[N,K] = dC[M,N]^T * A[M,K]          (wgrad of C = A W^T; contraction over the M rows, split over gridDim.z)
//   colsum  : out[n]     = sum_m in[m, n]                                   (bias gradients)
//   groupsum: out[b, c]  = sum_{r < rows} in[b*rows + r, col0 + c]          (gradient of the GEMM's row-group bias = dk3)
// Partials of the split reductions are written to a workspace and summed in a fixed order (no float atomics).
#pragma once
#include "common.cuh"

namespace digat {

constexpr int kWgTile = 128, kWgBK = 8, kWgThreads = 256;

// part[z][N][K] = sum over the z-th slice of rows of dC^T A
__global__ void __launch_bounds__(kWgThreads)
gemm_wgrad_kernel(const float* __restrict__ dC, int lddc, const float* __restrict__ A, int lda,
                  float* __restrict__ part, int M, int N, int K, int rows_per_split) {
    __shared__ __align__(16) float Ds[2][kWgBK][kWgTile];
    __shared__ __align__(16) float As[2][kWgBK][kWgTile];
    const int tid = threadIdx.x;
    const int n0 = blockIdx.y * kWgTile, k0 = blockIdx.x * kWgTile;
    const int m_lo = blockIdx.z * rows_per_split, m_hi = min(M, m_lo + rows_per_split);
    const int tx = tid & 15, ty = tid >> 4;
    const int lrow = tid >> 5, lq = (tid & 31) * 4;            // loader: 8 rows x 32 float4
    float4 rd, ra;
    auto gload = [&](int m0) {
        const int m = m0 + lrow;
        rd = (m < m_hi && n0 + lq < N) ? *reinterpret_cast<const float4*>(dC + (size_t)m * lddc + n0 + lq)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        ra = (m < m_hi && k0 + lq < K) ? *reinterpret_cast<const float4*>(A + (size_t)m * lda + k0 + lq)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto sstore = [&](int buf) {
        *reinterpret_cast<float4*>(&Ds[buf][lrow][lq]) = rd;
        *reinterpret_cast<float4*>(&As[buf][lrow][lq]) = ra;
    };
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int steps = (m_hi - m_lo + kWgBK - 1) / kWgBK;
    if (steps > 0) {
        gload(m_lo);
        sstore(0);
    }
    __syncthreads();
    for (int s = 0; s < steps; ++s) {
        const int buf = s & 1;
        if (s + 1 < steps) gload(m_lo + (s + 1) * kWgBK);
#pragma unroll
        for (int kk = 0; kk < kWgBK; ++kk) {
            float d[8], a[8];
#pragma unroll
            for (int gI = 0; gI < 2; ++gI) {
                const float4 v = *reinterpret_cast<const float4*>(&Ds[buf][kk][ty * 4 + gI * 64]);
                d[gI * 4 + 0] = v.x; d[gI * 4 + 1] = v.y; d[gI * 4 + 2] = v.z; d[gI * 4 + 3] = v.w;
                const float4 w = *reinterpret_cast<const float4*>(&As[buf][kk][tx * 4 + gI * 64]);
                a[gI * 4 + 0] = w.x; a[gI * 4 + 1] = w.y; a[gI * 4 + 2] = w.z; a[gI * 4 + 3] = w.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(d[i], a[j], acc[i][j]);
        }
        if (s + 1 < steps) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }
    float* out = part + (size_t)blockIdx.z * N * K;
#pragma unroll
    for (int gi = 0; gi < 2; ++gi)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int nrow = n0 + ty * 4 + gi * 64 + ii;
            if (nrow >= N) continue;
#pragma unroll
            for (int gj = 0; gj < 2; ++gj) {
                const int kc = k0 + tx * 4 + gj * 64;
                if (kc < K)       // K % 4 == 0: a quad is entirely inside or outside
                    *reinterpret_cast<float4*>(out + (size_t)nrow * K + kc) =
                        make_float4(acc[gi * 4 + ii][gj * 4 + 0], acc[gi * 4 + ii][gj * 4 + 1],
                                    acc[gi * 4 + ii][gj * 4 + 2], acc[gi * 4 + ii][gj * 4 + 3]);
            }
        }
}

// out[i] = sum_s part[s][i]   (fixed order: deterministic)
__global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, int splits, int64_t len) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= len) return;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < splits; ++z) {
        const float4 v = *reinterpret_cast<const float4*>(part + (size_t)z * len + i);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(out + i) = s;
}

// part[z][n] = sum over the z-th slice of rows of in[m, n]
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ in, int ld, float* __restrict__ part, int M, int N, int rows_per_split) {
    __shared__ float4 s_acc[8][32];
    const int cq = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int col = (blockIdx.x * 32 + cq) * 4;
    const int m_lo = blockIdx.y * rows_per_split, m_hi = min(M, m_lo + rows_per_split);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < N)
        for (int m = m_lo + rl; m < m_hi; m += 8) {
            const float4 v = *reinterpret_cast<const float4*>(in + (size_t)m * ld + col);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    s_acc[rl][cq] = s;
    __syncthreads();
    if (rl == 0 && col < N) {
#pragma unroll
        for (int r = 1; r < 8; ++r) {
            const float4 v = s_acc[r][cq];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        *reinterpret_cast<float4*>(part + (size_t)blockIdx.y * N + col) = s;
    }
}

// out[b, c] = sum_{r < rows} in[(b*rows + r) * ld + col0 + c]
__global__ void groupsum_kernel(const float* __restrict__ in, int ld, float* __restrict__ out, int groups, int rows,
                                int col0, int cols) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int cq = cols >> 2;
    if (i >= (int64_t)groups * cq) return;
    const int b = (int)(i / cq), q = (int)(i % cq);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* base = in + (size_t)b * rows * ld + col0 + 4 * q;
    for (int r = 0; r < rows; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(base + (size_t)r * ld);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(out + (size_t)b * cols + 4 * q) = s;
}

// Slices of the M reduction rows.  Enough CTAs to cover the SMs twice when the output has few tiles (the context
// projections of a training step: a 400 x 400 weight gradient contracted over 320 rows is 16 tiles -- one slice would
// leave 132 SMs idle for 40 serial k-steps), at least one slice per 2048 rows, at least 8 rows per slice, at most 64.
inline int reduce_splits(int M, int64_t tiles) {
    int64_t s = (296 + tiles - 1) / tiles;
    const int64_t lo = (M + 2047) / 2048, hi = M / 8 < 1 ? 1 : (M / 8 > 64 ? 64 : M / 8);
    if (s < lo) s = lo;
    if (s > hi) s = hi;
    return (int)(s < 1 ? 1 : s);
}
inline int wgrad_splits(int M, int N, int K) {
    return reduce_splits(M, (int64_t)((N + kWgTile - 1) / kWgTile) * ((K + kWgTile - 1) / kWgTile));
}
inline int colsum_splits(int M, int N) { return reduce_splits(M, (N / 4 + 31) / 32); }

// workspace floats needed by launch_linear_wgrad (M, N, K) / launch_colsum (M, N, 1)
inline int64_t wgrad_workspace_floats(int M, int N, int K) {
    const int64_t a = (int64_t)wgrad_splits(M, N, K) * N * K, b = K == 1 ? (int64_t)colsum_splits(M, N) * N : 0;
    return a > b ? a : b;
}

inline int launch_linear_wgrad(const float* dC, int lddc, const float* A, int lda, float* dW, float* workspace,
                               int M, int N, int K, cudaStream_t st) {
    DIGAT_REQUIRE(dC && A && dW && workspace, "digat_linear_wgrad: null pointer");
    DIGAT_REQUIRE(M > 0 && N > 0 && K > 0 && (N & 3) == 0 && (K & 3) == 0 && (lddc & 3) == 0 && (lda & 3) == 0,
                  "digat_linear_wgrad: N, K, lddc, lda must be multiples of 4");
    DIGAT_REQUIRE(aligned16(dC) && aligned16(A) && aligned16(dW) && aligned16(workspace),
                  "digat_linear_wgrad: pointers must be 16-byte aligned");
    const int splits = wgrad_splits(M, N, K);
    const int rows = ((M + splits - 1) / splits + kWgBK - 1) / kWgBK * kWgBK;
    dim3 grid((K + kWgTile - 1) / kWgTile, (N + kWgTile - 1) / kWgTile, splits);
    gemm_wgrad_kernel<<<grid, kWgThreads, 0, st>>>(dC, lddc, A, lda, workspace, M, N, K, rows);
    const int64_t len = (int64_t)N * K;
    reduce_partials_kernel<<<(unsigned)((len / 4 + 255) / 256), 256, 0, st>>>(workspace, dW, splits, len);
    return check_launch("digat_linear_wgrad");
}

// out[n] = sum_m in[m, n] in ONE launch for up to a few thousand rows (bias gradients of the context projections, the per-warp
// partials of the layer backward): 128 row lanes x 8 column quads per CTA, every lane's loads independent, the lanes added in a
// fixed order (warp shuffles, then the 32 warps).  (Two launches -- slices, then their sum -- cost 16 us for a [320 x 400] matrix.)
constexpr int kColsumSmallRows = 4096;
constexpr int kColsumSmallThreads = 1024;             // 128 row lanes x 8 column quads
__global__ void __launch_bounds__(kColsumSmallThreads)
colsum_small_kernel(const float* __restrict__ in, int ld, float* __restrict__ out, int M, int N, int accumulate) {
    __shared__ float4 s_acc[kColsumSmallThreads / 32][8];
    const int cq = threadIdx.x & 7, rl = threadIdx.x >> 3, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = (blockIdx.x * 8 + cq) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < N) {
#pragma unroll 8
        for (int m = rl; m < M; m += kColsumSmallThreads / 8) {
            const float4 v = *reinterpret_cast<const float4*>(in + (size_t)m * ld + col);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    // the four row lanes of a warp, then the 32 warps, in a fixed order
    s.x += __shfl_xor_sync(0xffffffffu, s.x, 8);  s.y += __shfl_xor_sync(0xffffffffu, s.y, 8);
    s.z += __shfl_xor_sync(0xffffffffu, s.z, 8);  s.w += __shfl_xor_sync(0xffffffffu, s.w, 8);
    s.x += __shfl_xor_sync(0xffffffffu, s.x, 16); s.y += __shfl_xor_sync(0xffffffffu, s.y, 16);
    s.z += __shfl_xor_sync(0xffffffffu, s.z, 16); s.w += __shfl_xor_sync(0xffffffffu, s.w, 16);
    if (lane < 8) s_acc[warp][lane] = s;
    __syncthreads();
    if (threadIdx.x < 8 && col < N) {
        float4 t = s_acc[0][cq];
#pragma unroll
        for (int w = 1; w < kColsumSmallThreads / 32; ++w) {
            const float4 v = s_acc[w][cq];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        float4* dst = reinterpret_cast<float4*>(out + col);
        if (accumulate) {
            const float4 o = *dst;
            t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
        }
        *dst = t;
    }
}

inline int launch_colsum(const float* in, int ld, float* out, float* workspace, int M, int N, cudaStream_t st,
                         int accumulate = 0) {
    DIGAT_REQUIRE(in && out && workspace, "digat_colsum: null pointer");
    DIGAT_REQUIRE(M > 0 && N > 0 && (N & 3) == 0 && (ld & 3) == 0 && aligned16(in) && aligned16(out) && aligned16(workspace),
                  "digat_colsum: N, ld must be multiples of 4 and pointers 16-byte aligned");
    if (M <= kColsumSmallRows && N <= 65536) {
        colsum_small_kernel<<<(N / 4 + 7) / 8, kColsumSmallThreads, 0, st>>>(in, ld, out, M, N, accumulate);
        return check_launch("digat_colsum");
    }
    DIGAT_REQUIRE(!accumulate, "digat_colsum: accumulate is supported for at most %d rows", kColsumSmallRows);
    const int splits = colsum_splits(M, N);
    const int rows = (M + splits - 1) / splits;
    dim3 grid((N / 4 + 31) / 32, splits);
    colsum_kernel<<<grid, 256, 0, st>>>(in, ld, splits == 1 ? out : workspace, M, N, rows);     // one slice: no second pass
    if (splits > 1) reduce_partials_kernel<<<(unsigned)((N / 4 + 255) / 256), 256, 0, st>>>(workspace, out, splits, N);
    return check_launch("digat_colsum");
}

// out[c, r] = in[r, c] (32 x 32 tiles through shared memory, both sides coalesced); with out_lo the transposed matrix is
// written as its two TF32 planes (hi = rna_tf32(x), lo = rna_tf32(x - hi)) -- the operands of the split-K weight-gradient
// GEMM, whose contraction index (the rows of dC and A) must be the contiguous one.
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ in, int ld_in, float* __restrict__ out, float* __restrict__ out_lo, int ld_out,
                 int rows, int cols) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + tx;
        tile[ty + 8 * k][tx] = (r < rows && c < cols) ? in[(size_t)r * ld_in + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;
        if (c < cols && r < rows) {
            const float x = tile[tx][ty + 8 * k];
            if (out_lo != nullptr) {
                uint32_t h;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
                const float hf = __uint_as_float(h);
                uint32_t l;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hf));
                out[(size_t)c * ld_out + r] = hf;
                out_lo[(size_t)c * ld_out + r] = __uint_as_float(l);
            } else {
                out[(size_t)c * ld_out + r] = x;
            }
        }
    }
}

inline int launch_transpose(const float* in, int ld_in, float* out, float* out_lo, int ld_out, int rows, int cols,
                            cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return DIGAT_OK;
    DIGAT_REQUIRE(in && out, "digat_transpose_f32: null pointer");
    DIGAT_REQUIRE(ld_in >= cols && ld_out >= rows, "digat_transpose_f32: leading dimension too small");
    dim3 grid((rows + 31) / 32, (cols + 31) / 32);
    DIGAT_REQUIRE(grid.y <= 65535, "digat_transpose_f32: too many columns (%d)", cols);
    transpose_kernel<<<grid, 256, 0, st>>>(in, ld_in, out, out_lo, ld_out, rows, cols);
    return check_launch("digat_transpose_f32");
}

inline int launch_groupsum(const float* in, int ld, float* out, int groups, int rows, int col0, int cols, cudaStream_t st) {
    DIGAT_REQUIRE(in && out, "digat_groupsum: null pointer");
    DIGAT_REQUIRE(groups > 0 && rows > 0 && cols > 0 && (cols & 3) == 0 && (col0 & 3) == 0 && (ld & 3) == 0 &&
                  aligned16(in) && aligned16(out), "digat_groupsum: bad shape / alignment");
    const int64_t total = (int64_t)groups * (cols / 4);
    groupsum_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, ld, out, groups, rows, col0, cols);
    return check_launch("digat_groupsum");
}

}  // namespace digat
