"""What would a COMPACT P buy the edge-driven layer kernel?  Runs it on graphs of n = 68 (MIND-shaped, ~38 active rows) and on
synthetic graphs that contain only the ~n_act active rows (n = 40 / 48, same edge count): the second is the traffic and the
tile shape a compacted layout would give.  Usage: python tools/layer_bench_n.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from digat_b200 import _lib
if len(sys.argv) > 1:
    _lib.LIB_PATH = sys.argv[1]
from digat_b200.graphEncoders import graph_layer_fwd

dev = torch.device('cuda:0')
B, D = 4096, 400
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(P, a, adj, X, iters=10, **kw):
    graph_layer_fwd(P, a, adj, X, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph_layer_fwd(P, a, adj, X, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


g = torch.Generator(device=dev).manual_seed(0)
for n, edges in ((68, 300), (48, 300), (40, 300), (40, 150), (32, 300), (24, 150)):
    p = min(1.0, edges / float(n * n))
    adj = (torch.rand(B, n, n, device=dev, generator=g) < p) | torch.eye(n, dtype=torch.bool, device=dev)
    X = torch.randn(B, n, D, device=dev, generator=g)
    P = torch.randn(B * n, 3 * D, device=dev, generator=g)
    a = torch.randn(D, device=dev, generator=g) * 0.1
    t = bench(P, a, adj, X)
    byt = B * (5 * n * D * 4 + n * n)
    print('n=%d edges/graph %.0f : %.4f ms  %.0f GB/s of its own bytes (%.2f GB)' % (n, adj.sum().item() / B, t, byt / t / 1e6, byt / 1e9))
