"""N = 400 GEMM (featureAffine shape): 128-wide tiles (variant 2) vs 208-wide persistent tiles (variant 0)."""
import torch, sys
sys.path.insert(0, '.')
from digat_b200 import _lib
def split(W):
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    _lib.call('digat_split_tf32', W.data_ptr(), hi.data_ptr(), lo.data_ptr(), W.numel(), 0); return hi, lo
for (M, N, K) in [(77824, 400, 400), (77824 + 77, 400, 400), (40960, 400, 400), (20000, 400, 800), (77824, 624, 400)]:
    g = torch.Generator().manual_seed(1)
    A = torch.randn(M, K, generator=g).cuda(); W = (torch.randn(N, K, generator=g) * 0.05).cuda(); b = torch.randn(N, generator=g).cuda()
    hi, lo = split(W)
    outs = {}
    for variant in (2, 0):
        _lib.call('digat_debug_set_gemm_variant', variant)
        C = torch.full((M, N + 16), 7.0, device='cuda')          # guard columns: must stay untouched
        def run(): _lib.call('digat_linear_tf32x3', A.data_ptr(), K, hi.data_ptr(), lo.data_ptr(), K, b.data_ptr(), C.data_ptr(), N + 16, M, N, K, 0, 1, 0, 0, 0, 0, 0)
        run(); torch.cuda.synchronize()
        assert bool((C[:, N:] == 7.0).all()), 'wrote past N'
        ref = A[:512].double() @ W.double().t() + b.double()
        err = float((C[:512, :N].double() - ref).abs().max() / ref.abs().max())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): run()
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        outs[variant] = C[:, :N].clone()
        print('variant', variant, (M, N, K), 'relerr %.2e' % err, 'ms %.4f TF %.1f' % (ms, 2 * M * N * K / ms / 1e9))
    print('  max |v0 - v2| / max|ref| = %.2e' % float((outs[0] - outs[2]).abs().max() / outs[2].abs().max()))
_lib.call('digat_debug_set_gemm_variant', 0)
