import torch, numpy as np, sys, time
sys.path.insert(0,'.')
from digat_b200 import _lib
def split(W):
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    _lib.call('digat_split_tf32', W.data_ptr(), hi.data_ptr(), lo.data_ptr(), W.numel(), 0); return hi, lo
st=0
for variant in (4,0):
    _lib.call('digat_debug_set_gemm_variant', variant)
    for (M,N,K) in [(16500,240,64),(20000,1200,400),(40960,1200,400),(278528,1200,400)]:
        g=torch.Generator().manual_seed(1)
        A=torch.randn(M,K,generator=g).cuda(); W=(torch.randn(N,K,generator=g)*0.05).cuda(); b=torch.randn(N,generator=g).cuda()
        hi,lo=split(W); C=torch.empty(M,N,device='cuda')
        def run(): _lib.call('digat_linear_tf32x3', A.data_ptr(), K, hi.data_ptr(), lo.data_ptr(), K, b.data_ptr(), C.data_ptr(), N, M, N, K, 0, 1, 0, 0, 0, 0, 0)
        run(); torch.cuda.synchronize()
        rows=torch.arange(0,M,max(1,M//256))[:256].cuda()
        ref=A[rows].double()@W.double().t()+b.double()
        err=float((C[rows].double()-ref).abs().max()/ref.abs().max())
        merr=float((C[rows].double()-ref).mean()/ref.abs().max())
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        for _ in range(3): run()
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/10
        print('variant',variant,(M,N,K),'relerr %.2e meanerr %.2e'%(err,merr),'ms %.3f TF(fp32-equiv) %.1f'%(ms, 2*M*N*K/ms/1e9))
