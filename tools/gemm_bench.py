"""Times the projection GEMM variants on the hot shapes (run on the GPU box): independent CTAs vs W multicast, 3xTF32 vs
TF32 + BF16 corrections.  Prints fp32-equivalent TFLOP/s (2MNK / CUDA-event time) and the error vs fp64 on sampled rows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from digat_b200 import _lib

def run(M, N, K, bf16c, cluster, reps=20):
    g = torch.Generator().manual_seed(1)
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) * 0.05).cuda()
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    st = torch.cuda.current_stream().cuda_stream
    _lib.call('digat_split_tf32', W.data_ptr(), hi.data_ptr(), lo.data_ptr(), W.numel(), st)
    hb = torch.empty(W.shape, dtype=torch.bfloat16, device='cuda'); lb = torch.empty_like(hb)
    _lib.call('digat_split_bf16', W.data_ptr(), hb.data_ptr(), lb.data_ptr(), W.numel(), st)
    C = torch.empty((M, N), device='cuda')
    _lib.call('digat_debug_set_gemm_variant', 6 + cluster)
    def call():
        if bf16c:
            _lib.call('digat_linear_tf32_bf16c', A.data_ptr(), K, hi.data_ptr(), hb.data_ptr(), lb.data_ptr(), K, 0, C.data_ptr(), N,
                      M, N, K, 0, 1, 0, 0, 0, 0, st)
        else:
            _lib.call('digat_linear_tf32x3', A.data_ptr(), K, hi.data_ptr(), lo.data_ptr(), K, 0, C.data_ptr(), N, M, N, K, 0, 1, 0, 0, 0, 0, st)
    for _ in range(3): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    sel = torch.randperm(M, generator=g)[:128]
    ref = A[sel].double() @ W.double().t()
    err = float((C[sel].double() - ref).abs().max() / ref.abs().max())
    print('M=%d N=%d K=%d  %-8s %-10s  %.3f ms  %.1f TFLOP/s  err %.2e' % (M, N, K, 'bf16c' if bf16c else 'tf32x3',
          ('single', 'multicast', '2-CTA MMA')[cluster], ms, 2.0 * M * N * K / ms / 1e9, err))

for (M, N, K) in [(136000, 1200, 400), (278528, 1200, 400), (27000, 1200, 400), (77824, 400, 400), (40960, 400, 800)]:
    for bf16c in (False, True):
        for cluster in (0, 2):
            run(M, N, K, bf16c, cluster)
_lib.call('digat_debug_set_gemm_variant', 6)
