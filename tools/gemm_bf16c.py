"""TF32+BF16-correction GEMM (digat_linear_tf32_bf16c) vs the 3xTF32 GEMM: error against fp64 and throughput."""
import sys
import torch
sys.path.insert(0, '.')
from digat_b200 import _lib

def planes(W):
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    _lib.call('digat_split_tf32', W.data_ptr(), hi.data_ptr(), lo.data_ptr(), W.numel(), 0)
    hb = torch.empty(W.shape, dtype=torch.bfloat16, device=W.device); lb = torch.empty_like(hb)
    _lib.call('digat_split_bf16', W.data_ptr(), hb.data_ptr(), lb.data_ptr(), W.numel(), 0)
    return hi, lo, hb, lb

for (M, N, K, scale) in [(40960, 1200, 400, 1.0), (278528, 1200, 400, 1.0), (154000, 1200, 400, 1.0), (77824, 400, 400, 1.0),
                         (40960, 400, 800, 1.0), (40960, 1200, 400, 30.0)]:
    g = torch.Generator().manual_seed(1)
    A = (torch.randn(M, K, generator=g) * scale).cuda()
    if scale != 1.0:
        A = torch.relu(A)                                  # non-negative activations with a large mean: worst case for sums
    W = (torch.randn(N, K, generator=g) * 0.05).cuda(); b = torch.randn(N, generator=g).cuda()
    hi, lo, hb, lb = planes(W)
    rows = torch.arange(0, M, max(1, M // 512))[:512].cuda()
    ref = A[rows].double() @ W.double().t() + b.double()
    out = {}
    for name in ('3xtf32', 'tf32+bf16'):
        C = torch.full((M, N), 7.0, device='cuda')
        if name == '3xtf32':
            run = lambda: _lib.call('digat_linear_tf32x3', A.data_ptr(), K, hi.data_ptr(), lo.data_ptr(), K, b.data_ptr(), C.data_ptr(), N, M, N, K, 0, 1, 0, 0, 0, 0, 0)
        else:
            run = lambda: _lib.call('digat_linear_tf32_bf16c', A.data_ptr(), K, hi.data_ptr(), hb.data_ptr(), lb.data_ptr(), K, b.data_ptr(), C.data_ptr(), N, M, N, K, 0, 1, 0, 0, 0, 0, 0)
        run(); torch.cuda.synchronize()
        d = (C[rows].double() - ref)
        err, rms = float(d.abs().max() / ref.abs().max()), float(d.pow(2).mean().sqrt() / ref.abs().max())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): run()
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out[name] = C
        print('%-10s %s scale %.0f  max err %.2e rms %.2e  %.4f ms  %.1f TFLOP/s fp32-equiv' % (name, (M, N, K), scale, err, rms, ms, 2 * M * N * K / ms / 1e9))
    print('   max |diff| between schemes / max|ref| = %.2e' % float((out['3xtf32'] - out['tf32+bf16']).abs().max() / ref.abs().max()))
