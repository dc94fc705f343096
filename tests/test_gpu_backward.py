"""GPU tests of the training path: the backward kernels against torch autograd of the CPU oracle (fp32 parity target,
fp64 ground truth), dropout off (train mode with p=0, the only setting in which the reference is reproducible,
SURVEY.md 7.2 item 6) plus a mask-injection test of the attention-weight dropout."""
import numpy as np
import pytest
import torch

from oracle import digat_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
ORDER = ('news_graph_embeddings', 'news_graph', 'news_graph_mask', 'user_news_embedding', 'user_graph',
         'user_category_mask', 'user_category_indices')
GRAD_TOL = 2e-5     # gradients vs the fp64 oracle autograd, relative to the largest magnitude of the tensor


def _batch(cfg, rows, seed, scale=0.3):
    from digat_b200 import synth
    corpus = synth.make_corpus(cfg, n_news=400, n_behaviors=max(4, rows // 2), mean_candidates=3.0, seed=seed, emb_scale=scale)
    ids = np.arange(min(rows, corpus.pair_behavior.shape[0]))
    return synth.make_batch(corpus, ids)


def _oracle_grads(sd, batch, dtype, wn, wu):
    P = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    b = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in batch.items()}
    b['news_graph_embeddings'] = b['news_graph_embeddings'].clone().requires_grad_(True)
    b['user_news_embedding'] = b['user_news_embedding'].clone().requires_grad_(True)
    cn, cu = O.forward(P, *[b[k] for k in ORDER])
    loss = (cn * wn.to(dtype)).sum() + (cu * wu.to(dtype)).sum() + (cn * cu).sum()
    loss.backward()
    g = {k: v.grad for k, v in P.items()}
    g['in:news'] = b['news_graph_embeddings'].grad
    g['in:hist'] = b['user_news_embedding'].grad
    return g, float(loss)


@pytest.mark.parametrize('N,L,rows,edge_driven', [(3, 2, 12, True), (5, 1, 6, True), (3, 2, 12, False)])
def test_encoder_gradients_match_oracle(N, L, rows, edge_driven, monkeypatch):
    from digat_b200 import autograd_ops
    monkeypatch.setattr(autograd_ops, 'TRAIN_EDGE_DRIVEN', edge_driven)
    from digat_b200 import synth
    from digat_b200.graphEncoders import DIGAT
    cfg = synth.make_config(SAG_neighbors=N, SAG_hops=2, graph_depth=L, dropout_rate=0.0)
    sd = synth.make_state_dict(cfg, seed=4)
    batch = _batch(cfg, rows, seed=N)
    B = batch['news_graph'].shape[0]
    g = torch.Generator().manual_seed(1)
    wn, wu = torch.randn(B, 400, generator=g), torch.randn(B, 400, generator=g)
    ref32, l32 = _oracle_grads(sd, batch, torch.float32, wn, wu)
    ref64, l64 = _oracle_grads(sd, batch, torch.float64, wn, wu)

    m = DIGAT(cfg, 400)
    m.load_state_dict(sd)
    m = m.cuda().train()                       # train mode, p = 0
    b = {k: v.cuda() for k, v in batch.items()}
    b['news_graph_embeddings'].requires_grad_(True)
    b['user_news_embedding'].requires_grad_(True)
    cn, cu = m(*[b[k] for k in ORDER])
    loss = (cn * wn.cuda()).sum() + (cu * wu.cuda()).sum() + (cn * cu).sum()
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - l64) / abs(l64) < 1e-5
    ours = {k: v.grad for k, v in m.named_parameters()}
    ours['in:news'] = b['news_graph_embeddings'].grad
    ours['in:hist'] = b['user_news_embedding'].grad
    assert set(ours) == set(ref32)
    worst, bad = 0.0, []
    for k in sorted(ref32):
        assert ours[k] is not None, 'no gradient for ' + k
        e64 = rel_err(ours[k].cpu().numpy(), ref64[k].numpy())
        eref = rel_err(ref32[k].numpy(), ref64[k].numpy())
        worst = max(worst, e64)
        print('grad %-45s ours vs fp64 %.2e   oracle fp32 vs fp64 %.2e' % (k, e64, eref))
        if e64 >= GRAD_TOL:
            bad.append((k, e64, eref))
    print('worst gradient rel err vs fp64: %.3e' % worst)
    assert not bad, 'gradients beyond %.0e of the fp64 oracle: %s' % (GRAD_TOL, bad)


@pytest.mark.parametrize('edge_driven,n,D,density,p', [(False, 26, 400, 0.4, 0.2), (True, 26, 400, 0.4, 0.2),
                                                        (True, 68, 400, 0.12, 0.2), (True, 68, 400, 1.0, 0.0),
                                                        (True, 7, 36, 0.3, 0.5), (True, 33, 200, 0.2, 0.0)])
def test_graph_layer_dropout_mask_injection(edge_driven, n, D, density, p):
    """alpha~ = alpha * keep / (1-p) inside the fused kernel, forward and backward, against plain torch with the SAME mask;
    dense [B,n,n] kernels and the edge-driven pair (CSR forward with per-edge outputs + digat_graph_layer_bwd_csr)."""
    from digat_b200 import _lib
    from digat_b200.autograd_ops import GraphLayerFn
    from digat_b200.graphEncoders import build_graph_csr
    g = torch.Generator().manual_seed(3)
    B = 5
    P = (torch.randn(B * n, 3 * D, generator=g) * 0.5)
    a = torch.randn(D, generator=g) * 0.1
    X = torch.randn(B, n, D, generator=g)
    adj = (torch.rand(B, n, n, generator=g) < density) | torch.eye(n, dtype=torch.bool)
    adj[1, n // 2, :] = False                    # an edge-less row: uniform weights, no gradient into the scores
    adj[2, :, 0] = False                         # a node nobody attends to: dh, dU of that node are exactly zero
    keep = (torch.rand(B, n, n, generator=g) >= p) if p > 0 else None
    dY = torch.randn(B, n, D, generator=g)

    def ref(P, a, X, dtype):
        P, a, X = P.to(dtype).requires_grad_(True), a.to(dtype).requires_grad_(True), X.to(dtype).requires_grad_(True)
        h, U, K2 = P.view(B, n, 3 * D).split(D, dim=2)
        s = (torch.relu(U.unsqueeze(1) + K2.unsqueeze(2)) * a).sum(-1)
        e = torch.nn.functional.leaky_relu(s, 0.2)
        al = torch.softmax(e.masked_fill(adj == 0, -1e9), dim=2)
        if keep is not None:
            al = al * keep.to(dtype) / (1 - p)
        Y = torch.relu(torch.bmm(al, h)) + X
        Y.backward(dY.to(dtype))
        return Y.detach(), P.grad, a.grad, X.grad

    Y64, dP64, da64, dX64 = ref(P, a, X, torch.float64)
    Pc, ac, Xc = P.cuda().requires_grad_(True), a.cuda().requires_grad_(True), X.cuda().requires_grad_(True)
    csr = None
    if edge_driven:
        assert _lib.load().digat_graph_layer_csr_training_supported(n, D) == 1
        csr = build_graph_csr(adj.cuda(), transpose=True)
    Y = GraphLayerFn.apply(Pc, ac, adj.cuda(), Xc, None if keep is None else keep.cuda(), 1.0 / (1 - p), csr)
    Y.backward(dY.cuda())
    torch.cuda.synchronize()
    if edge_driven:
        assert float(Pc.grad.view(B, n, 3 * D)[2, 0, :2 * D].abs().max()) == 0.0
    assert rel_err(Y.detach().cpu().numpy(), Y64.numpy()) < 1e-5
    assert rel_err(Pc.grad.cpu().numpy(), dP64.numpy()) < 1e-4
    assert rel_err(ac.grad.cpu().numpy(), da64.numpy()) < 1e-4
    assert rel_err(Xc.grad.cpu().numpy(), dX64.numpy()) < 1e-6


def test_pool_and_segment_backward():
    from digat_b200.autograd_ops import AttentionPoolFn, TopicSegmentFn
    from oracle.scatter_shim import scatter_softmax, scatter_sum
    g = torch.Generator().manual_seed(8)
    B, m, D, H, S = 7, 19, 400, 50, 19
    Fm = torch.randn(B, m, D, generator=g)
    T = torch.randn(B, m, D, generator=g)
    v = torch.randn(B, D, generator=g) * 0.3
    mask = torch.rand(B, m, generator=g) < 0.7
    mask[0] = False
    dout = torch.randn(B, D, generator=g)
    for use_resid in (False, True):
        F64, T64, v64 = (t.double().requires_grad_(True) for t in (Fm, T, v))
        Fp = torch.relu(F64) + T64 if use_resid else F64
        al = torch.softmax(((Fp * v64.unsqueeze(1)).sum(-1) / 20.0).masked_fill(mask == 0, -1e9), 1)
        out = (al.unsqueeze(2) * Fp).sum(1)
        out.backward(dout.double())
        Fc, Tc, vc = (t.cuda().requires_grad_(True) for t in (Fm, T, v))
        o = AttentionPoolFn.apply(Fc, vc, mask.cuda(), Tc if use_resid else None)
        o.backward(dout.cuda())
        torch.cuda.synchronize()
        assert rel_err(o.detach().cpu().numpy(), out.detach().numpy()) < 1e-5
        assert rel_err(Fc.grad.cpu().numpy(), F64.grad.numpy()) < 1e-5
        assert rel_err(vc.grad.cpu().numpy(), v64.grad.numpy()) < 1e-5
        if use_resid:
            assert rel_err(Tc.grad.cpu().numpy(), T64.grad.numpy()) < 1e-5
    Xu = torch.randn(B, H + 18, D, generator=g)
    cidx = torch.randint(0, S, (B, H), generator=g)
    cidx[1] = 18
    dT = torch.randn(B, S, D, generator=g)
    X64, v64 = Xu.double().requires_grad_(True), v.double().requires_grad_(True)
    Xh = X64[:, :H]
    al = scatter_softmax((Xh * v64.unsqueeze(1)).sum(-1) / 20.0, cidx, 1, dim_size=S)
    Tt = scatter_sum(al.unsqueeze(2) * Xh, cidx, dim=1, dim_size=S)
    Tt.backward(dT.double())
    Xc, vc = Xu.cuda().requires_grad_(True), v.cuda().requires_grad_(True)
    err = torch.zeros(1, dtype=torch.int32, device='cuda')
    Tg = TopicSegmentFn.apply(Xc, vc, cidx.cuda(), H, S, err)
    Tg.backward(dT.cuda())
    torch.cuda.synchronize()
    assert rel_err(Tg.detach().cpu().numpy(), Tt.detach().numpy()) < 1e-5
    assert rel_err(Xc.grad.cpu().numpy(), X64.grad.numpy()) < 1e-5
    assert rel_err(vc.grad.cpu().numpy(), v64.grad.numpy()) < 1e-5
    assert float(Xc.grad[:, H:].abs().max()) == 0.0


@pytest.mark.parametrize('M,N,K,rows', [(680, 1200, 400, 68), (130, 400, 400, 10), (5000, 400, 800, 1)])
def test_linear_backward(M, N, K, rows):
    from digat_b200.autograd_ops import lin
    g = torch.Generator().manual_seed(M)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) * 0.05
    bias = torch.randn(N, generator=g)
    gb = torch.randn(M // rows, 400, generator=g)
    dC = torch.randn(M, N, generator=g)
    A64, W64, b64, g64 = (t.double().requires_grad_(True) for t in (A, W, bias, gb))
    C = A64 @ W64.t() + b64
    C = torch.cat([C[:, :0], C[:, 0:400] + g64.repeat_interleave(rows, 0), C[:, 400:]], 1)
    C.backward(dC.double())
    Ac, Wc, bc, gc = (t.cuda().requires_grad_(True) for t in (A, W, bias, gb))
    out = lin(Ac, Wc, bc, group_bias=gc, group_rows=rows, group_col0=0)
    out.backward(dC.cuda())
    torch.cuda.synchronize()
    assert rel_err(out.detach().cpu().numpy(), C.detach().numpy()) < 5e-6
    assert rel_err(Ac.grad.cpu().numpy(), A64.grad.numpy()) < 5e-6
    assert rel_err(Wc.grad.cpu().numpy(), W64.grad.numpy()) < 2e-6
    assert rel_err(bc.grad.cpu().numpy(), b64.grad.numpy()) < 2e-6
    assert rel_err(gc.grad.cpu().numpy(), g64.grad.numpy()) < 2e-6


def test_train_mode_with_dropout_runs_and_is_stochastic():
    from digat_b200 import synth
    from digat_b200.graphEncoders import DIGAT
    cfg = synth.make_config(graph_depth=2, dropout_rate=0.2)
    sd = synth.make_state_dict(cfg, seed=2)
    batch = _batch(cfg, 8, seed=9)
    m = DIGAT(cfg, 400)
    m.load_state_dict(sd)
    m = m.cuda().train()
    b = [batch[k].cuda() for k in ORDER]
    torch.manual_seed(0)
    cn1, cu1 = m(*b)
    cn2, cu2 = m(*b)
    (cn1 * cu1).sum().backward()
    torch.cuda.synchronize()
    assert torch.isfinite(cn1).all() and torch.isfinite(cu1).all()
    assert not torch.equal(cn1, cn2)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    m.eval()
    with torch.no_grad():
        e1, _ = m(*b)
        e2, _ = m(*b)
    assert torch.equal(e1, e2)


@pytest.mark.parametrize('n,D,p', [(68, 400, 0.2), (13, 400, 0.0), (7, 36, 0.5)])
def test_edge_driven_backward_extras_match_the_plain_call(n, D, p):
    """digat_graph_layer_bwd_csr with the in-kernel relu mask and the per-graph column sums: dP bit-identical to the call that
    gets G = dY * mask precomputed, dh_sum / du_sum = the column sums of the dh / dU blocks of dP per graph, and the whole
    call repeatable bit for bit (no atomics)."""
    from digat_b200 import _lib
    from digat_b200.graphEncoders import _ptr, _stream, build_graph_csr, graph_layer_fwd
    g = torch.Generator().manual_seed(n)
    B = 6
    P = (torch.randn(B * n, 3 * D, generator=g) * 0.5).cuda()
    a = (torch.randn(D, generator=g) * 0.1).cuda()
    X = torch.randn(B, n, D, generator=g).cuda()
    adj = ((torch.rand(B, n, n, generator=g) < 0.2) | torch.eye(n, dtype=torch.bool)).cuda()
    keep = (torch.rand(B, n, n, generator=g) >= p).cuda() if p > 0 else None
    scale = 1.0 / (1 - p) if p > 0 else 1.0
    dY = torch.randn(B, n, D, generator=g).cuda()
    csr = build_graph_csr(adj, transpose=True)
    score = torch.empty(B, n * n, device='cuda')
    alpha = torch.empty(B, n * n, device='cuda')
    rmask = torch.empty(B, n, D, device='cuda', dtype=torch.uint8)
    graph_layer_fwd(P, a, adj, X, drop_keep=keep, drop_scale=scale, score_out=score, alpha_out=alpha, relu_mask_out=rmask,
                    csr=(csr[0], csr[1], None))

    def run(G, mask, sums):
        parts = _lib.load().digat_graph_layer_bwd_csr_parts()
        dP = torch.empty_like(P)
        da = torch.empty(B * parts, D, device='cuda')
        dh = torch.empty(B * parts, D, device='cuda') if sums else None
        du = torch.empty(B * parts, D, device='cuda') if sums else None
        _lib.call('digat_graph_layer_bwd_csr', P.data_ptr(), P.stride(0), a.data_ptr(), csr[0].data_ptr(), csr[1].data_ptr(),
                  csr[2].data_ptr(), csr[3].data_ptr(), score.data_ptr(), alpha.data_ptr(), _ptr(keep), scale, G.data_ptr(),
                  _ptr(mask), dP.data_ptr(), dP.stride(0), da.data_ptr(), _ptr(dh), _ptr(du), B, n, D, _stream())
        torch.cuda.synchronize()
        return dP, da, dh, du
    dP0, da0, _, _ = run((dY * rmask).contiguous(), None, False)
    dP1, da1, dh, du = run(dY, rmask, True)
    dP2, da2, dh2, du2 = run(dY, rmask, True)
    assert torch.equal(dP0, dP1) and torch.equal(da0, da1)
    assert torch.equal(dP1, dP2) and torch.equal(da1, da2) and torch.equal(dh, dh2) and torch.equal(du, du2)
    blocks = dP1.view(B, n, 3 * D).double()
    for got, want in ((dh, blocks[:, :, :D].sum(1)), (du, blocks[:, :, D:2 * D].sum(1))):      # per-warp partial rows
        assert rel_err(got.view(B, -1, D).double().sum(1).cpu().numpy(), want.cpu().numpy()) < 1e-6
