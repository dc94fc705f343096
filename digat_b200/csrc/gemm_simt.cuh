// Exact-fp32 CUDA-core GEMM:  C[M,N] = A[M,K] * W[N,K]^T (+bias) (relu).   Both operands K-major (nn.Linear layout).
// Used for the skinny context projections (M = batch) and as the fp32 ground truth the tcgen05 3xTF32 GEMM is
// tested against.  128x128x8 CTA tile, 256 threads, 8x8 register micro-tile, double-buffered smem.
#pragma once
#include "common.cuh"

namespace digat {

// Row-group bias: C[m, col0 + c] += ptr[(m / rows) * cols + c] for c in [0, cols).  Used to fold the per-graph
// context projection k3 into the K1 block of the node projections: U = fl(k3 + K1), the first broadcast add of
// Eq. (8) (reference graphEncoders.py:150), with the reference's rounding.
struct GroupBias {
    const float* ptr;
    int rows, col0, cols, ld;      // ld = row pitch of ptr in elements (>= cols)
    // Row scatter (tcgen05 persistent GEMM only): product row m is written to row out_rows[m] of C and takes the
    // row-group bias of THAT row.  A holds only the rows worth computing (digat_b200/graphEncoders.py: nodes whose
    // output can reach a context), C keeps the dense [graphs * n] layout the fused layer kernel streams with TMA.
    const int32_t* out_rows = nullptr;
    // Split-K batches (persistent GEMM only; weight gradients contract over tens of thousands of rows): the contraction
    // range is cut into kbatches slices, slice s writes its partial product to C + s * c_batch_stride (summed by the
    // caller with exact fp32 adds).  More tiles than SMs for skinny outputs, and at most K/kbatches truncating
    // accumulate steps per accumulator.
    int kbatches = 1;
    long long c_batch_stride = 0;
};

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_tn_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                   const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int N, int K, int relu,
                   GroupBias gb) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int KQ = BK / 4;               // float4 per tile row
    static_assert((BM * BK / 4) % NT == 0 && (BN * BK / 4) % NT == 0, "tile/loader mismatch");
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Ws[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);

    constexpr int A_LD = BM * BK / 4 / NT;   // float4 loads per thread per k-tile
    constexpr int W_LD = BN * BK / 4 / NT;
    float4 ra[A_LD], rw[W_LD];

    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
            int f = tid + i * NT;
            int row = f / KQ, kq = (f % KQ) * 4;
            int gm = m0 + row;
            ra[i] = (gm < M && k0 + kq < K) ? *reinterpret_cast<const float4*>(A + (size_t)gm * lda + k0 + kq)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < W_LD; ++i) {
            int f = tid + i * NT;
            int row = f / KQ, kq = (f % KQ) * 4;
            int gn = n0 + row;
            rw[i] = (gn < N && k0 + kq < K) ? *reinterpret_cast<const float4*>(W + (size_t)gn * ldw + k0 + kq)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
            int f = tid + i * NT;
            int row = f / KQ, kq = (f % KQ) * 4;
            As[buf][kq + 0][row] = ra[i].x; As[buf][kq + 1][row] = ra[i].y;
            As[buf][kq + 2][row] = ra[i].z; As[buf][kq + 3][row] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < W_LD; ++i) {
            int f = tid + i * NT;
            int row = f / KQ, kq = (f % KQ) * 4;
            Ws[buf][kq + 0][row] = rw[i].x; Ws[buf][kq + 1][row] = rw[i].y;
            Ws[buf][kq + 2][row] = rw[i].z; Ws[buf][kq + 3][row] = rw[i].w;
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nk = (K + BK - 1) / BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], w[TN];
            // rows owned by a thread are strided by 4-float groups: (ty*4 + g*(BM/2)) keeps float4 smem reads conflict-free
#pragma unroll
            for (int g = 0; g < TM / 4; ++g) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4 + g * (BM / (TM / 4))]);
                a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int g = 0; g < TN / 4; ++g) {
                float4 v = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4 + g * (BN / (TN / 4))]);
                w[g * 4 + 0] = v.x; w[g * 4 + 1] = v.y; w[g * 4 + 2] = v.z; w[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }

#pragma unroll
    for (int gi = 0; gi < TM / 4; ++gi)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int gm = m0 + ty * 4 + gi * (BM / (TM / 4)) + ii;
            if (gm >= M) continue;
#pragma unroll
            for (int gj = 0; gj < TN / 4; ++gj) {
                const int gn = n0 + tx * 4 + gj * (BN / (TN / 4));
                float o[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    float v = acc[gi * 4 + ii][gj * 4 + jj];
                    if (bias != nullptr && gn + jj < N) v += bias[gn + jj];
                    if (gb.ptr != nullptr && gn + jj >= gb.col0 && gn + jj < gb.col0 + gb.cols)
                        v += gb.ptr[(size_t)(gm / gb.rows) * gb.ld + (gn + jj - gb.col0)];
                    if (relu) v = fmaxf(v, 0.f);
                    o[jj] = v;
                }
                float* dst = C + (size_t)gm * ldc + gn;
                if (gn + 3 < N && ((ldc & 3) == 0)) {
                    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        if (gn + jj < N) dst[jj] = o[jj];
                }
            }
        }
}

inline int launch_linear_f32(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc,
                             int M, int N, int K, int relu, GroupBias gb, cudaStream_t st) {
    if (M == 0) return DIGAT_OK;                        // empty batch: nothing to do (pointers of empty tensors are null)
    DIGAT_REQUIRE(A && W && C, "digat_linear_f32: null pointer");
    DIGAT_REQUIRE(M >= 0 && N > 0 && K > 0, "digat_linear_f32: bad shape M=%d N=%d K=%d", M, N, K);
    DIGAT_REQUIRE((K & 3) == 0 && (lda & 3) == 0 && (ldw & 3) == 0, "digat_linear_f32: K, lda, ldw must be multiples of 4");
    DIGAT_REQUIRE(aligned16(A) && aligned16(W) && aligned16(C), "digat_linear_f32: pointers must be 16-byte aligned");
    DIGAT_REQUIRE(lda >= K && ldw >= K && ldc >= N, "digat_linear_f32: leading dimension too small");
    DIGAT_REQUIRE(gb.ptr == nullptr || (gb.rows > 0 && gb.col0 >= 0 && gb.cols > 0 && gb.col0 + gb.cols <= N && gb.ld >= gb.cols),
                  "digat_linear_f32: bad row-group bias");
    if (M == 0) return DIGAT_OK;
    if (M <= 2048) {
        // skinny: smaller tiles so that the grid covers the 148 SMs
        constexpr int BM = 64, BN = 64;
        dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
        gemm_tn_f32_kernel<BM, BN, 16, 4, 4><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, C, ldc, M, N, K, relu, gb);
    } else {
        constexpr int BM = 128, BN = 128;
        dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
        gemm_tn_f32_kernel<BM, BN, 8, 8, 8><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, C, ldc, M, N, K, relu, gb);
    }
    return check_launch("digat_linear_f32");
}

}  // namespace digat
