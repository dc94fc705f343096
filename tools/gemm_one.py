import sys, torch
sys.path.insert(0, '.')
from digat_b200 import _lib
variant = int(sys.argv[1]); M, N, K = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
_lib.call('digat_debug_set_gemm_variant', variant)
g = torch.Generator().manual_seed(1)
A = torch.randn(M, K, generator=g).cuda(); W = (torch.randn(N, K, generator=g) * 0.05).cuda()
hi, lo = torch.empty_like(W), torch.empty_like(W)
_lib.call('digat_split_tf32', W.data_ptr(), hi.data_ptr(), lo.data_ptr(), W.numel(), 0)
C = torch.empty(M, N, device='cuda')
for _ in range(3):
    _lib.call('digat_linear_tf32x3', A.data_ptr(), K, hi.data_ptr(), lo.data_ptr(), K, 0, C.data_ptr(), N, M, N, K, 0, 1, 0, 0, 0, 0, 0)
torch.cuda.synchronize()
