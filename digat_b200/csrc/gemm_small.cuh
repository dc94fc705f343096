// Exact-fp32 CUDA-core GEMM for the few-hundred-row projections of a training step (context queries / keys, k3, gates:
// M = batch rows).  On the tcgen05 path such a product is a dozen 128-row tiles whose 3 x K/8 serial MMAs cost ~9 us
// whatever the tile width (measured: 690 cycles per 16-wide k-block), plus a TF32 split of the weight, plus a
// transposed copy for dgrad and two for wgrad.  Here one kernel covers the three products without any operand copy:
//   out[i, j] = sum_c L(i, c) * R(c, j) (+ bias[j])
//   forward  C  = A W^T  : L = A  [i = m, c = k contiguous]            R = W  stored [j = n][c = k]  (r_trans)
//   dgrad    dA = dC W   : L = dC [i = m, c = n contiguous]            R = W  stored [c = n][j = k]
//   wgrad    dW = dC^T A : L = dC stored [c = m][i = n]   (l_trans)    R = A  stored [c = m][j = k]
// 32 x 32 output tile per CTA, 32-deep contraction chunks double-buffered in shared memory in c-major order.
// A [320 x 400] output is 130 CTAs: one per SM.  Fixed summation order: deterministic.
#pragma once
#include "common.cuh"

namespace digat {

constexpr int kSgT = 32, kSgC = 64, kSgGroups = 8, kSgThreads = 32 * kSgGroups;
constexpr int kSgPitch = kSgT + 4;   // tile row pitch: float4-aligned rows
constexpr int kSgCg = kSgC / kSgGroups;   // contraction steps of a chunk per warp
constexpr int kSgAhead = 2;          // chunks in flight from global memory (registers): 2 x 16 KB per CTA
constexpr int kSgLoaders = kSgThreads / 2;         // threads per operand
constexpr int kSgLoads = kSgC * kSgT / 4 / kSgLoaders;   // float4 per loader thread, operand and chunk

// One WARP computes the whole 32 x 32 tile (8 x 4 outputs per lane) for its share of the contraction: warp g takes the
// c-steps [8g, 8g + 8) of every 64-deep chunk, and the eight partial tiles are added by a fixed pairwise tree at the end
// (deterministic).  (Sixteen warps, four per scheduler, were slower: 10.6 vs 9.5 us for [320 x 400 x 400].)
// Why this shape: the product is bound by shared-memory reads, not FMAs -- with 2 x 4 outputs per thread every c-step cost
// 6 LSU cycles per warp for 8 FMA instructions (measured 33 cycles per c-step and tile, 4x the FMA time); 8 x 4 outputs
// read 12 floats for 32 FMAs.  It is also latency-bound from global memory (a chunk's math takes ~0.1 us, a load ~0.7 us):
// each thread keeps its float4s of the next four chunks in registers.
template <bool L_T, bool R_T>
__global__ void __launch_bounds__(kSgThreads)
gemm_small_f32_kernel(const float* __restrict__ L, int ldl, const float* __restrict__ R, int ldr,
                      const float* __restrict__ bias, float* __restrict__ out, int ldo, int I, int J, int C) {
    __shared__ __align__(16) float tiles[2 * 2 * kSgC * kSgPitch];        // L and R chunks, double-buffered; then the partial tiles
    static_assert(sizeof(float4) * (kSgGroups / 2) * 8 * 32 <= sizeof(float) * 2 * 2 * kSgC * kSgPitch, "partials must fit the tiles");
    float (*Ls)[kSgC][kSgPitch] = reinterpret_cast<float (*)[kSgC][kSgPitch]>(tiles);
    float (*Rs)[kSgC][kSgPitch] = reinterpret_cast<float (*)[kSgC][kSgPitch]>(tiles + 2 * kSgC * kSgPitch);
    float4 (*red)[8][32] = reinterpret_cast<float4 (*)[8][32]>(tiles);
    const int tid = threadIdx.x, grp = tid >> 5, lane = tid & 31;
    const int i0 = blockIdx.y * kSgT, j0 = blockIdx.x * kSgT;
    const int ti = lane >> 3, tj = lane & 7;               // outputs (i0 + 8 ti + {0..7}, j0 + 4 tj + {0..3})
    // loaders: the first half of the CTA owns L's 512 float4 of a chunk (four each), the second half R's.  An operand whose contiguous
    // dimension is c is transposed on the way into shared memory: lanes take 32 different rows, so the four scalar stores
    // of a float4 hit 32 different banks; the other layout is copied quad for quad.
    const bool isR = tid >= kSgLoaders;
    const int e = tid % kSgLoaders;
    const bool transposing = isR ? R_T : !L_T;
    // element t of a thread: (row, quad) of the [32 rows x kSgC] (c contiguous, transposed on store) or [kSgC x 32] tile
    int lr[kSgLoads], lq[kSgLoads];
#pragma unroll
    for (int t = 0; t < kSgLoads; ++t) {
        const int f = e + t * kSgLoaders;
        lr[t] = transposing ? (f & 31) : (f >> 3);
        lq[t] = transposing ? (f >> 5) * 4 : (f & 7) * 4;
    }
    float4 rv[kSgAhead][kSgLoads];

    // Loads are UNCONDITIONAL (clamped addresses) and zeroed at store time when out of range: a guarded load compiles to a
    // branch that waits for the load it skips over, which made every prefetch synchronous (22 % of the stall samples sat on
    // those branches and on the first store behind them).
    const float* base = isR ? R : L;
    const int ld = isR ? ldr : ldl;
    const int fixed0 = isR ? j0 : i0, fixedN = isR ? J : I;       // the tile's own dimension (rows of L / columns of R)
    auto in_range = [&](int t, int c0) {
        const int r = lr[t], q = lq[t];
        return transposing ? (fixed0 + r < fixedN && c0 + q < C) : (c0 + r < C && fixed0 + q < fixedN);
    };
    auto gload = [&](float4 (&v)[kSgLoads], int c0) {
#pragma unroll
        for (int t = 0; t < kSgLoads; ++t) {
            const int r = lr[t], q = lq[t];
            // transposing: stored [tile dim][c] -> row = fixed0 + r, quad along c;   else stored [c][tile dim]
            const int row = transposing ? min(fixed0 + r, fixedN - 1) : min(c0 + r, C - 1);
            const int col = transposing ? min(c0 + q, (C - 4) & ~3) : min(fixed0 + q, fixedN - 4);
            v[t] = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * ld + col));
        }
    };
    auto sstore = [&](int buf, const float4 (&v)[kSgLoads], int c0) {
        float (*dst)[kSgPitch] = isR ? Rs[buf] : Ls[buf];
#pragma unroll
        for (int t = 0; t < kSgLoads; ++t) {
            const int r = lr[t], q = lq[t];
            const float4 x = in_range(t, c0) ? v[t] : make_float4(0.f, 0.f, 0.f, 0.f);
            if (transposing) {
                dst[q + 0][r] = x.x; dst[q + 1][r] = x.y; dst[q + 2][r] = x.z; dst[q + 3][r] = x.w;
            } else {
                *reinterpret_cast<float4*>(&dst[r][q]) = x;
            }
        }
    };

    float acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const int chunks = (C + kSgC - 1) / kSgC;
#pragma unroll
    for (int t = 0; t < kSgAhead; ++t) gload(rv[t], t * kSgC);           // (beyond C: clamped reads, never stored)
    sstore(0, rv[0], 0);
    __syncthreads();
    gload(rv[0], kSgAhead * kSgC);
    for (int ch0 = 0; ch0 < chunks; ch0 += kSgAhead) {
#pragma unroll
        for (int u = 0; u < kSgAhead; ++u) {
            const int ch = ch0 + u, buf = u & 1;                          // kSgAhead is even: chunk parity = u parity
            if (ch < chunks) {
                // the operands of four c-steps are fetched together: one shared-memory round trip per 128 FMAs
#pragma unroll
                for (int cb = 0; cb < kSgCg; cb += 4) {
                    float4 la[4], lb[4], rr[4];
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const int c = grp * kSgCg + cb + w;
                        la[w] = *reinterpret_cast<const float4*>(&Ls[buf][c][8 * ti]);
                        lb[w] = *reinterpret_cast<const float4*>(&Ls[buf][c][8 * ti + 4]);
                        rr[w] = *reinterpret_cast<const float4*>(&Rs[buf][c][4 * tj]);
                    }
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const float lv[8] = {la[w].x, la[w].y, la[w].z, la[w].w, lb[w].x, lb[w].y, lb[w].z, lb[w].w};
#pragma unroll
                        for (int a = 0; a < 8; ++a) {
                            acc[a][0] = fmaf(lv[a], rr[w].x, acc[a][0]); acc[a][1] = fmaf(lv[a], rr[w].y, acc[a][1]);
                            acc[a][2] = fmaf(lv[a], rr[w].z, acc[a][2]); acc[a][3] = fmaf(lv[a], rr[w].w, acc[a][3]);
                        }
                    }
                }
                if (ch + 1 < chunks) {
                    sstore(buf ^ 1, rv[(u + 1) % kSgAhead], (ch + 1) * kSgC);   // chunk ch + 1, loaded kSgAhead chunks ago
                    __syncthreads();
                    gload(rv[(u + 1) % kSgAhead], (ch + 1 + kSgAhead) * kSgC);
                }
            }
        }
    }
    // the tiles are dead: their memory now holds partial tiles.  Pairwise tree in a fixed order: warps [h, 2h) park, warps
    // [0, h) add their partner, h = 4, 2, 1 (deterministic).
#pragma unroll
    for (int h = kSgGroups / 2; h >= 1; h >>= 1) {
        __syncthreads();
        if (grp >= h && grp < 2 * h) {
#pragma unroll
            for (int a = 0; a < 8; ++a) red[grp - h][a][lane] = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
        }
        __syncthreads();
        if (grp < h) {
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const float4 v = red[grp][a][lane];
                acc[a][0] += v.x; acc[a][1] += v.y; acc[a][2] += v.z; acc[a][3] += v.w;
            }
        }
    }
    if (grp > 0) return;
    const int j = j0 + 4 * tj;
    if (j < J) {                     // J % 4 == 0: a quad is entirely inside or outside
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias != nullptr) bv = *reinterpret_cast<const float4*>(bias + j);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int i = i0 + 8 * ti + a;
            if (i < I)
                *reinterpret_cast<float4*>(out + (size_t)i * ldo + j) =
                    make_float4(acc[a][0] + bv.x, acc[a][1] + bv.y, acc[a][2] + bv.z, acc[a][3] + bv.w);
        }
    }
}

inline int launch_gemm_f32_small(const float* L, int ldl, int l_trans, const float* R, int ldr, int r_trans, const float* bias,
                                 float* out, int ldo, int I, int J, int C, cudaStream_t st) {
    if (I == 0 || J == 0) return DIGAT_OK;
    DIGAT_REQUIRE(L && R && out, "digat_gemm_f32_small: null pointer");
    DIGAT_REQUIRE(I > 0 && J > 0 && C > 0 && (J & 3) == 0 && (ldl & 3) == 0 && (ldr & 3) == 0 && (ldo & 3) == 0 && ldo >= J,
                  "digat_gemm_f32_small: J and the leading dimensions must be multiples of 4");
    DIGAT_REQUIRE(l_trans ? ((I & 3) == 0 && ldl >= I) : ((C & 3) == 0 && ldl >= C),
                  "digat_gemm_f32_small: the contiguous dimension of L must be a multiple of 4");
    DIGAT_REQUIRE(r_trans ? ((C & 3) == 0 && ldr >= C) : ldr >= J, "digat_gemm_f32_small: bad R layout");
    DIGAT_REQUIRE(!(l_trans && r_trans), "digat_gemm_f32_small: l_trans and r_trans together are not a product of this path");
    DIGAT_REQUIRE(aligned16(L) && aligned16(R) && aligned16(out) && (!bias || aligned16(bias)),
                  "digat_gemm_f32_small: pointers must be 16-byte aligned");
    dim3 grid((J + kSgT - 1) / kSgT, (I + kSgT - 1) / kSgT);
    if (l_trans) gemm_small_f32_kernel<true, false><<<grid, kSgThreads, 0, st>>>(L, ldl, R, ldr, bias, out, ldo, I, J, C);
    else if (r_trans) gemm_small_f32_kernel<false, true><<<grid, kSgThreads, 0, st>>>(L, ldl, R, ldr, bias, out, ldo, I, J, C);
    else gemm_small_f32_kernel<false, false><<<grid, kSgThreads, 0, st>>>(L, ldl, R, ldr, bias, out, ldo, I, J, C);
    return check_launch("digat_gemm_f32_small");
}

}  // namespace digat
