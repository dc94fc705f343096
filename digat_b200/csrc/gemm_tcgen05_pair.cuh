// 2-CTA (cta_group::2) variant of the tcgen05 3xTF32 projection GEMM for large M.
//
// Two CTAs of one cluster (a TPC pair of SMs) cooperate on a 256 x BN output tile: each CTA loads and splits ITS 128
// rows of A and only HALF of the W tile (BN/2 rows of both TF32 planes); the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256), which reads A and W from both CTAs' shared memory and writes each CTA's 128
// accumulator rows into that CTA's TMEM.  Per flop this halves the W bytes every SM has to pull through L2/TMA -- the
// single-CTA kernel (gemm_tcgen05.cuh) is bound by exactly that traffic (measured: 185 TFLOP/s with one half-tile of W
// traffic per 128 rows, 204 with two row blocks sharing W, see tools/gemm_exp.py) -- and leaves room for 6 stages.
//
// Synchronisation (s = pipeline stage; "L" = lives in the leader CTA only):
//   full[s]        per CTA : own TMA bytes landed (A rows + W half)            -> own transform warps
//   ready[s]   L           : 4 local transform warps + 1 remote arrive forwarded by the peer's idle MMA-warp thread
//   peer_ready[s] (peer)   : the peer's 4 transform warps (cta scope)     -> the peer's forwarder thread
//   empty[s]       per CTA : tcgen05.commit.multicast from the leader           -> each CTA's TMA producer
//   accum_full     per CTA : tcgen05.commit.multicast after the last k-block    -> each CTA's epilogue warps
#pragma once
#include "gemm_tcgen05.cuh"

namespace digat {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    __syncwarp();                                        // warps whose elected lane ran a role loop reconverge first
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
// arrive (release, cluster scope) on the mbarrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; ; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity), "r"(kMbarSuspendHintNs) : "memory");
        if (done) return;
        if (spin > (1u << 23)) {
            printf("digat: cluster mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {      // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

template <int BN>
struct TcPairCfg {
    static constexpr int A_BYTES = 128 * kTcBK * 4;                    // one TF32 plane of this CTA's 128 rows
    static constexpr int WH_ROWS = BN / 2;                             // rows of W this CTA loads
    static constexpr int W_SLOT = ((WH_ROWS * kTcBK * 4 + 1023) / 1024) * 1024;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_SLOT;
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
    static constexpr int TMEM_COLS = 2 * BN <= 256 ? 256 : 512;        // main + correction accumulators
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 512;
    static_assert(BN % 16 == 0 && BN <= 256 && (BN / 2) % 8 == 0, "invalid UMMA N for cta_group::2");
};

template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tf32x3_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_whi,
                        const __grid_constant__ CUtensorMap map_wlo, const float* __restrict__ bias,
                        float* __restrict__ C, int ldc, int M, int N, int K, GroupBias gb) {
    using Cfg = TcPairCfg<BN>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* ready = bars + Cfg::STAGES;
    uint64_t* empty = bars + 2 * Cfg::STAGES;
    uint64_t* accum_full = bars + 3 * Cfg::STAGES;
    uint64_t* peer_ready = bars + 3 * Cfg::STAGES + 1;   // [STAGES] peer CTA only: its 4 transform warps (cta scope)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * Cfg::STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();             // 0 = leader (rows 0..127 of the pair), 1 = peer (rows 128..255)
    // grid.x = 2 * (N tiles): the pair (cluster dims (2,1,1)) is adjacent along x and the N tiles of one 256-row block
    // are launched back to back, so the A rows are read from DRAM once and re-read from L2
    const int n0 = (int)(blockIdx.x >> 1) * BN, m0 = ((int)blockIdx.y * 2 + (int)rank) * 128;
    const int nkb = (K + kTcBK - 1) / kTcBK;
    __shared__ float bias_s[BN];
    for (int i = threadIdx.x; i < BN; i += kTcThreads) bias_s[i] = (bias != nullptr && n0 + i < N) ? bias[n0 + i] : 0.f;

    auto a_hi = [&](int s) { return smem + (size_t)s * Cfg::STAGE_BYTES; };
    auto a_lo = [&](int s) { return smem + (size_t)s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
    auto w_hi = [&](int s) { return smem + (size_t)s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES; };
    auto w_lo = [&](int s) { return smem + (size_t)s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES + Cfg::W_SLOT; };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&map_whi)) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&map_wlo)) : "memory");
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&ready[s], kTcTransformThreads / 32 + 1);  // leader: its 4 transform warps + 1 forwarded arrive of the peer
            mbar_init(&peer_ready[s], kTcTransformThreads / 32);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // both CTAs of the pair issue the 2-CTA allocation from the same logical warp
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                   // barriers of both CTAs are initialised before any remote arrive
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (each CTA: its A rows, its half of W)
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % Cfg::STAGES;
                const uint32_t ph = (kb / Cfg::STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&full[s], Cfg::A_BYTES + 2 * Cfg::WH_ROWS * kTcBK * 4);
                tma_load_2d(a_hi(s), &map_a, &full[s], kb * kTcBK, m0);
                tma_load_2d(w_hi(s), &map_whi, &full[s], kb * kTcBK, n0 + (int)rank * Cfg::WH_ROWS);
                tma_load_2d(w_lo(s), &map_wlo, &full[s], kb * kTcBK, n0 + (int)rank * Cfg::WH_ROWS);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer: leader CTA only
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(256, BN);
            const uint32_t d_main = tmem_base, d_corr = tmem_base + (uint32_t)BN;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % Cfg::STAGES;
                const uint32_t ph = (kb / Cfg::STAGES) & 1;
                mbar_wait_cluster(&ready[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t d_ahi = umma_desc_sw64(smem_u32(a_hi(s))), d_alo = umma_desc_sw64(smem_u32(a_lo(s)));
                const uint64_t d_whi = umma_desc_sw64(smem_u32(w_hi(s))), d_wlo = umma_desc_sw64(smem_u32(w_lo(s)));
#pragma unroll
                for (int k = 0; k < kTcBK / 8; ++k) {
                    const uint64_t koff = (uint64_t)((k * 8 * 4) >> 4);
                    const uint32_t first = (kb > 0 || k > 0) ? 1u : 0u;
                    umma_tf32_2cta(d_main, d_ahi + koff, d_whi + koff, idesc, first);
                    umma_tf32_2cta(d_corr, d_alo + koff, d_whi + koff, idesc, first);
                    umma_tf32_2cta(d_corr, d_ahi + koff, d_wlo + koff, idesc, 1u);
                }
                umma_commit_2cta(&empty[s]);
            }
            umma_commit_2cta(accum_full);
        } else if (rank == 1 && lane == 0) {
            // forwarder: cluster-scope releases cost a MEMBAR each; keep them off the transform warps -- this otherwise
            // idle thread turns "the peer's 4 transform warps are done with stage s" into ONE remote arrive
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % Cfg::STAGES;
                mbar_wait(&peer_ready[s], (kb / Cfg::STAGES) & 1);
                mbar_arrive_cluster(&ready[s], 0);
            }
        }
    } else {
        // ------------------------------------------------------------------ operand transform (both CTAs), then epilogue
        const int t = threadIdx.x - 64;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % Cfg::STAGES;
            const uint32_t ph = (kb / Cfg::STAGES) & 1;
            mbar_wait(&full[s], ph);
            float4* hi = reinterpret_cast<float4*>(a_hi(s));
            float4* lo = reinterpret_cast<float4*>(a_lo(s));
#pragma unroll
            for (int i = 0; i < Cfg::A_BYTES / 16 / kTcTransformThreads; ++i) {
                const int idx = t + i * kTcTransformThreads;
                const float4 v = hi[idx];
                float4 vh, vl;
                vh.x = to_tf32_rna(v.x); vh.y = to_tf32_rna(v.y); vh.z = to_tf32_rna(v.z); vh.w = to_tf32_rna(v.w);
                vl.x = to_tf32_rna(v.x - vh.x); vl.y = to_tf32_rna(v.y - vh.y);
                vl.z = to_tf32_rna(v.z - vh.z); vl.w = to_tf32_rna(v.w - vh.w);
                hi[idx] = vh;
                lo[idx] = vl;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // each SM reads its own A rows: CTA-local visibility suffices
            __syncwarp();
            if (lane == 0) mbar_arrive(rank == 0 ? &ready[s] : &peer_ready[s]);   // cta-scope arrive, one per warp
        }
        mbar_wait(accum_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        const bool has_gb = gb.ptr != nullptr;
        const int m = m0 + q * 32 + lane;
        const bool row_ok = m < M;
        float* crow = C + (size_t)m * ldc + n0;
        const float* grow = (has_gb && row_ok) ? gb.ptr + (size_t)(m / gb.rows) * gb.ld - gb.col0 + n0 : nullptr;
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t rm[16], rc[16];
        float4 gq[4];
        auto issue_group = [&](int c) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(rm[0]), "=r"(rm[1]), "=r"(rm[2]), "=r"(rm[3]), "=r"(rm[4]), "=r"(rm[5]), "=r"(rm[6]), "=r"(rm[7]),
                  "=r"(rm[8]), "=r"(rm[9]), "=r"(rm[10]), "=r"(rm[11]), "=r"(rm[12]), "=r"(rm[13]), "=r"(rm[14]), "=r"(rm[15])
                : "r"(tbase + (uint32_t)c));
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(rc[0]), "=r"(rc[1]), "=r"(rc[2]), "=r"(rc[3]), "=r"(rc[4]), "=r"(rc[5]), "=r"(rc[6]), "=r"(rc[7]),
                  "=r"(rc[8]), "=r"(rc[9]), "=r"(rc[10]), "=r"(rc[11]), "=r"(rc[12]), "=r"(rc[13]), "=r"(rc[14]), "=r"(rc[15])
                : "r"(tbase + (uint32_t)(BN + c)));
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
                const int col = n0 + c + v4 * 4;
                gq[v4] = (grow != nullptr && col >= gb.col0 && col < gb.col0 + gb.cols)
                             ? *reinterpret_cast<const float4*>(grow + c + v4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        issue_group(0);
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = (__uint_as_float(rm[i]) + __uint_as_float(rc[i])) + bias_s[c + i];
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
                o[v4 * 4 + 0] += gq[v4].x; o[v4 * 4 + 1] += gq[v4].y;
                o[v4 * 4 + 2] += gq[v4].z; o[v4 * 4 + 3] += gq[v4].w;
            }
            if (c + 16 < BN) issue_group(c + 16);
            if (row_ok) {
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4)
                    if (n0 + c + v4 * 4 < N)
                        *reinterpret_cast<float4*>(crow + c + v4 * 4) =
                            make_float4(o[v4 * 4 + 0], o[v4 * 4 + 1], o[v4 * 4 + 2], o[v4 * 4 + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                   // neither CTA frees TMEM / exits while the pair is still working
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;"
                     :: "r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

template <int BN>
int launch_tf32x3_pair(const float* A, int lda, const float* W_hi, const float* W_lo, int ldw, const float* bias,
                              float* C, int ldc, int M, int N, int K, GroupBias gb, cudaStream_t st) {
    using Cfg = TcPairCfg<BN>;
    CUtensorMap ma, mh, ml;
    int rc;
    if ((rc = make_tensor_map_2d(&ma, A, M, K, lda, 128, kTcBK, CU_TENSOR_MAP_SWIZZLE_64B)) != DIGAT_OK) return rc;
    if ((rc = make_tensor_map_2d(&mh, W_hi, N, K, ldw, Cfg::WH_ROWS, kTcBK, CU_TENSOR_MAP_SWIZZLE_64B)) != DIGAT_OK) return rc;
    if ((rc = make_tensor_map_2d(&ml, W_lo, N, K, ldw, Cfg::WH_ROWS, kTcBK, CU_TENSOR_MAP_SWIZZLE_64B)) != DIGAT_OK) return rc;
    if (int rc_ = ensure_dynamic_smem(gemm_tf32x3_pair_kernel<BN>, (size_t)(Cfg::SMEM))) return rc_;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * ((N + BN - 1) / BN), (M + 255) / 256);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DIGAT_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32x3_pair_kernel<BN>, ma, mh, ml, bias, C, ldc, M, N, K, gb));
    return check_launch("digat_linear_tf32x3(pair)");
}

}  // namespace digat
