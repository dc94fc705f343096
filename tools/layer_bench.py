"""Micro-benchmark of the edge-driven layer kernel on MIND-shaped user graphs (B = 4096 graphs of 68 nodes).
Usage: python tools/layer_bench.py [path/to/alternate/libdigat_sm100.so]   (e.g. a -DDIGAT_TC_TIMING build)"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from digat_b200 import _lib
if len(sys.argv) > 1:
    _lib.LIB_PATH = sys.argv[1]
from digat_b200 import synth
from digat_b200.graphEncoders import graph_layer_fwd

B, D = 4096, 400
cfg = synth.make_config(SAG_neighbors=3, SAG_hops=2, graph_depth=3)
corpus = synth.make_corpus(cfg, n_news=3000, n_behaviors=400, mean_candidates=12.0, seed=3)
beh = corpus.pair_behavior[:B].astype(np.int64)
dev = torch.device('cuda:0')
adj = torch.from_numpy(corpus.user_graph[beh]).to(dev)
cidx = torch.from_numpy(corpus.user_category_indices[beh]).to(dev)
cmask = torch.from_numpy(corpus.user_category_mask[beh]).to(dev)
n = adj.shape[1]
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn(B, n, D, device=dev, generator=g)
P = torch.randn(B * n, 3 * D, device=dev, generator=g)
a = torch.randn(D, device=dev, generator=g) * 0.1
act = torch.empty((B, n), dtype=torch.uint8, device=dev)
_lib.call('digat_user_active_rows', adj.data_ptr(), 0, cidx.data_ptr(), cmask.data_ptr(), act.data_ptr(), 0, B, n, 50, 19, 0)
torch.cuda.synchronize()
print('B=%d n=%d edges/graph %.0f active rows/graph %.1f active edges/graph %.0f' % (
    B, n, adj.sum().item() / B, act.sum().item() / B, (adj & (act[:, :, None] != 0)).sum().item() / B))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


from digat_b200.graphEncoders import build_graph_csr


def bench(row_active, iters=10, csr=None):
    if csr is not None:
        return bench_csr(row_active, iters, csr)
    Y = graph_layer_fwd(P, a, adj, X, row_active=row_active)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph_layer_fwd(P, a, adj, X, row_active=row_active)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return Y, float(np.median(ts))


def bench_csr(row_active, iters, csr):
    Y = graph_layer_fwd(P, a, adj, X, row_active=row_active, csr=csr)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph_layer_fwd(P, a, adj, X, row_active=row_active, csr=csr)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return Y, float(np.median(ts))


Y0, t0 = bench(None)
Y1, t1 = bench(act)
alg = B * (5 * n * D * 4 + n * n + D * 4)
print('all rows    : %.4f ms  %.0f GB/s algorithmic' % (t0, alg / t0 / 1e6))
print('active rows : %.4f ms  %.0f GB/s algorithmic' % (t1, alg / t1 / 1e6))
k = act != 0
print('active rows identical:', bool(torch.equal(Y0[k], Y1[k])))
csr = build_graph_csr(adj, act) + (None,)
Y2, t2 = bench(act, csr=csr)
print('active rows, precomputed CSR : %.4f ms  %.0f GB/s algorithmic; identical: %s' % (t2, alg / t2 / 1e6, bool(torch.equal(Y2[k], Y1[k]))))
_lib.call('digat_debug_set_layer_mode', 1)
Yd = graph_layer_fwd(P, a, adj, X)
_lib.call('digat_debug_set_layer_mode', 0)
print('max |sparse - dense| = %.3e' % float((Y0 - Yd).abs().max()))


# ---- news graphs (n = 10): dense multi-graph kernel vs edge-driven kernel with several graphs per CTA
nid = corpus.pair_news[:B].astype(np.int64)
adj_n = torch.from_numpy(corpus.news_graph[nid]).to(dev)
nn_ = adj_n.shape[1]
Xn = torch.randn(B, nn_, D, device=dev, generator=g)
Pn = torch.randn(B * nn_, 3 * D, device=dev, generator=g)
res = {}
for mode, name in ((1, 'dense'), (0, 'edge-driven')):
    _lib.call('digat_debug_set_layer_mode', mode)
    Yn = graph_layer_fwd(Pn, a, adj_n, Xn)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph_layer_fwd(Pn, a, adj_n, Xn); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res[name] = Yn
    algn = B * (5 * nn_ * D * 4 + nn_ * nn_ + D * 4)
    print('news n=%d %-12s: %.4f ms  %.0f GB/s algorithmic' % (nn_, name, float(np.median(ts)), algn / float(np.median(ts)) / 1e6))
_lib.call('digat_debug_set_layer_mode', 0)
print('news max |sparse - dense| = %.3e' % float((res['dense'] - res['edge-driven']).abs().max()))
