"""Node pruning of the user graph at inference (digat_user_active_rows + row-scattered projection GEMM + row_active in
the edge-driven layer kernel): flags against a literal numpy statement of the rule, GEMM scatter bit-identical to the
full GEMM, and encoder outputs bit-identical with pruning on and off."""
import numpy as np
import pytest
import torch

from tests.helpers import case_inputs

pytestmark = pytest.mark.gpu
ORDER = ('news_graph_embeddings', 'news_graph', 'news_graph_mask', 'user_news_embedding', 'user_graph',
         'user_category_mask', 'user_category_indices')


def _active_rule(adj, cidx, cmask, H):
    """Literal statement of the pruning rule (include/digat_sm100.h, digat_user_active_rows)."""
    G, n, _ = adj.shape
    out = np.zeros((G, n), dtype=np.uint8)
    for g in range(G):
        a = adj[g]
        if (~a.any(axis=1)).any():                      # an edge-less row: uniform softmax over all nodes
            out[g] = 1
            continue
        off = a & ~np.eye(n, dtype=bool)
        col_used = off.any(axis=0)
        any_bucket = cmask[g].any()
        for i in range(n):
            pooled = i < H and (not any_bucket or cmask[g, cidx[g, i]])
            out[g, i] = col_used[i] or pooled
    return out


def test_active_rows_kernel_matches_rule():
    from digat_b200 import _lib, graphs
    rng = np.random.Generator(np.random.PCG64(21))
    H, C, G = 50, 18, 200
    lens = rng.integers(0, H + 1, size=G)
    lens[:4] = [0, H, 1, 2]
    cats = rng.integers(0, C, size=(G, H))
    adj, cmask, cidx = graphs.build_user_graphs(cats, lens, H, C)
    adj[5, 60, :] = False                                # an edge-less row -> keep everything
    adj[6, 3, 40] = True                                 # asymmetric edge: node 3 attends to padded node 40 -> 40 is active
    cmask[7, :] = False                                  # all buckets masked -> every history slot is pooled
    cmask[8, C] = True                                   # padding bucket unmasked -> padded slots are pooled
    dev = torch.device('cuda:0')
    t = lambda x: torch.from_numpy(x).to(dev)
    act = torch.empty((G, H + C), dtype=torch.uint8, device=dev)
    adj_d, cidx_d, cmask_d = t(adj), t(cidx), t(cmask)            # keep the device tensors alive across the calls
    pooled = torch.empty_like(act)
    _lib.call('digat_user_active_rows', adj_d.data_ptr(), 0, cidx_d.data_ptr(), cmask_d.data_ptr(), act.data_ptr(),
              pooled.data_ptr(), G, H + C, H, C + 1, torch.cuda.current_stream().cuda_stream)
    want = _active_rule(adj, cidx, cmask, H)
    assert np.array_equal(act.cpu().numpy(), want)
    pl = pooled.cpu().numpy()                              # pooled rows: history slots of visible buckets; subset of active
    assert not pl[:, H:].any() and (pl <= want).all()
    for g_ in (0, 1, 2, 3, 7, 8, 20):
        vis = cmask[g_].any()
        assert np.array_equal(pl[g_, :H], np.array([(not vis) or cmask[g_, cidx[g_, i]] for i in range(H)], dtype=np.uint8))
    assert want[5].all() and want[6, 40] == 1 and want[7, :H].all() and want[8, :H].all()
    assert 0.2 < want.mean() < 0.8                       # MIND-shaped graphs: roughly half of the nodes are prunable
    # indexed form: graphs and segment ids per behaviour, masks per pair
    idx = torch.from_numpy(rng.integers(0, G, size=333).astype(np.int32)).to(dev)
    act2 = torch.empty((333, H + C), dtype=torch.uint8, device=dev)
    cm2 = cmask_d[idx.long()].contiguous()
    _lib.call('digat_user_active_rows', adj_d.data_ptr(), idx.data_ptr(), cidx_d.data_ptr(), cm2.data_ptr(),
              act2.data_ptr(), 0, 333, H + C, H, C + 1, torch.cuda.current_stream().cuda_stream)
    assert np.array_equal(act2.cpu().numpy(), want[idx.cpu().numpy()])


@pytest.mark.parametrize('M_full,N,frac', [(68 * 700, 1200, 0.5), (68 * 300, 1200, 0.05), (19 * 500, 400, 0.7), (68 * 40, 1200, 0.5)])
def test_gemm_row_scatter_bit_identical(M_full, N, frac):
    from digat_b200.graphEncoders import PackedWeight, linear
    g = torch.Generator().manual_seed(M_full + N)
    K, rows_per_group = 400, 68 if M_full % 68 == 0 else 19
    X = torch.randn(M_full, K, generator=g).cuda()
    W = PackedWeight((torch.randn(N, K, generator=g) * 0.05).cuda())
    bias = torch.randn(N, generator=g).cuda()
    gbias = torch.randn(M_full // rows_per_group, 400, generator=g).cuda()
    full = torch.empty(M_full + 300, N, device='cuda')[:M_full]              # M > 256 -> tensor-core path
    linear(X, W, bias, out=full, group_bias=gbias, group_rows=rows_per_group, group_col0=0 if N == 400 else 400)
    keep = torch.rand(M_full, generator=g) < frac
    keep[:rows_per_group] = False                                            # a whole group without rows
    rows = keep.nonzero().squeeze(1).to(torch.int32).cuda()
    A = X[rows.long()].contiguous()
    out = torch.full((M_full, N), 7.0, device='cuda')
    linear(A, W, bias, out=out, group_bias=gbias, group_rows=rows_per_group, group_col0=0 if N == 400 else 400, c_rows=rows)
    torch.cuda.synchronize()
    k = keep.cuda()
    assert torch.equal(out[k], full[k])                                      # same MMAs per row: bit-identical
    assert bool((out[~k] == 7.0).all())                                      # unlisted rows untouched


def _encoder(name):
    from digat_b200.graphEncoders import DIGAT
    cfg, sd, corpus, batch = case_inputs(name)
    m = DIGAT(cfg, 400)
    m.load_state_dict(sd)
    return m.cuda().eval(), {k: v.cuda() for k, v in batch.items()}


@pytest.mark.parametrize('name', ['default_n3_L3', 'code_default_n5_L2', 'wide_n8_L7'])
def test_inference_identical_with_and_without_pruning(name):
    m, b = _encoder(name)
    from digat_b200 import _lib
    args = [b[k] for k in ORDER]
    c0 = m.compute_news_graph_context(b['news_graph_embeddings'], b['news_graph_mask'])
    m.prune_user_nodes = True
    before = _lib.launch_count()
    cn1, cu1 = m.inference(*args, c0)
    pruned_launches = _lib.launch_count() - before
    m.prune_user_nodes = False
    before = _lib.launch_count()
    cn0, cu0 = m.inference(*args, c0)
    if args[0].shape[0] * 68 >= 256:                                         # smaller batches stay un-pruned (exact-fp32 GEMM)
        assert pruned_launches > _lib.launch_count() - before                # the pruned path really ran (extra gathers)
    assert torch.equal(cn1, cn0) and torch.equal(cu1, cu0)
    fn1, fu1 = m.forward(*args)
    m.prune_user_nodes = True
    fn2, fu2 = m.forward(*args)
    assert torch.equal(fn1, fn2) and torch.equal(fu1, fu2)


def test_scorer_identical_with_and_without_pruning():
    from tests.test_gpu_scoring import _setup
    cfg, sd, corpus, scorer = _setup(n_beh=50, seed=9)
    beh = torch.from_numpy(corpus.pair_behavior).cuda()
    news = torch.from_numpy(corpus.pair_news).cuda()
    scorer.enc.prune_user_nodes = True
    a = scorer.score_resident(beh, news)
    a2 = scorer.score_resident(beh, news, share_user_graphs=False)
    scorer.enc.prune_user_nodes = False
    b = scorer.score_resident(beh, news)
    assert torch.equal(a, b) and torch.equal(a2, b)


def test_news_active_rows_kernel_matches_rule():
    from digat_b200 import _lib, synth
    rng = np.random.Generator(np.random.PCG64(4))
    cfg = synth.make_config(SAG_neighbors=5, SAG_hops=2, graph_depth=1)
    node, adj, mask = synth.make_sag(rng, 400, cfg.news_graph_size, 5, 2)
    n = adj.shape[1]
    adj[7, 3, :] = False                                  # an edge-less row -> keep everything
    mask[8, :] = False                                    # every entry masked -> uniform softmax reads every node
    adj[9, 2, n - 1] = True                               # node 2 attends to the (possibly unused) last slot
    want = np.zeros((400, n), dtype=np.uint8)
    for g in range(400):
        if (~adj[g].any(axis=1)).any() or not mask[g].any():
            want[g] = 1
            continue
        col_used = (adj[g] & ~np.eye(n, dtype=bool)).any(axis=0)
        want[g] = col_used | mask[g] | (np.arange(n) == 0)
    dev = torch.device('cuda:0')
    adj_d, mask_d = torch.from_numpy(adj).to(dev), torch.from_numpy(mask).to(dev)
    act = torch.empty((400, n), dtype=torch.uint8, device=dev)
    _lib.call('digat_news_active_rows', adj_d.data_ptr(), mask_d.data_ptr(), act.data_ptr(), 400, n,
              torch.cuda.current_stream().cuda_stream)
    assert np.array_equal(act.cpu().numpy(), want)
    assert want[9, n - 1] == 1 and 0.2 < want.mean() < 0.9


def test_graph_csr_builder_matches_numpy():
    """digat_build_graph_csr: row pointers (bit 15 = edge-less row -> uniform), edge records neighbour | row << 8, pruned
    rows empty; records read through adj_index."""
    from digat_b200.graphEncoders import build_graph_csr
    rng = np.random.Generator(np.random.PCG64(12))
    G, n = 9, 37
    adj = rng.random((5, n, n)) < 0.15
    adj[:, np.arange(n), np.arange(n)] = True
    adj[2, 7, :] = False                                             # an edge-less row (cannot happen in reference data)
    act = rng.random((G, n)) < 0.7
    idx = rng.integers(0, 5, size=G).astype(np.int32)
    rowptr, meta = build_graph_csr(torch.from_numpy(adj).cuda(), torch.from_numpy(act.astype(np.uint8)).cuda(),
                                   torch.from_numpy(idx).cuda())
    torch.cuda.synchronize()
    rowptr, meta = rowptr.cpu().numpy().view(np.uint16), meta.cpu().numpy().view(np.uint16)
    for g in range(G):
        e = 0
        assert rowptr[g, 0] == 0
        for i in range(n):
            cols = np.nonzero(adj[idx[g], i])[0] if act[g, i] else np.zeros(0, dtype=np.int64)
            uniform = bool(act[g, i]) and len(cols) == 0
            if uniform:
                cols = np.arange(n)
            assert np.array_equal(meta[g, e:e + len(cols)], (cols | (i << 8)).astype(np.uint16)), (g, i)
            e += len(cols)
            assert rowptr[g, i + 1] == (e | (0x8000 if uniform else 0)), (g, i)


def test_graph_csr_transpose_lists_incoming_edges():
    """colptr / cedge of digat_build_graph_csr: column j's range holds the CSR positions of the edges (i, j), rows ascending
    (what digat_graph_layer_bwd_csr walks for dh_j and dU_j)."""
    from digat_b200.graphEncoders import build_graph_csr
    rng = np.random.Generator(np.random.PCG64(13))
    for G, n, dens in ((6, 68, 0.1), (3, 128, 0.3), (4, 5, 0.5), (2, 33, 1.0)):
        adj = rng.random((G, n, n)) < dens
        adj[:, np.arange(n), np.arange(n)] = True
        adj[0, n // 2, :] = False                                    # edge-less row: listed with every node (uniform softmax)
        if n > 2:
            adj[G - 1, :, 1] = False                                 # a node without incoming edges
        rowptr, meta, colptr, cedge = build_graph_csr(torch.from_numpy(adj).cuda(), transpose=True)
        torch.cuda.synchronize()
        rowptr, meta = rowptr.cpu().numpy().view(np.uint16), meta.cpu().numpy().view(np.uint16)
        colptr, cedge = colptr.cpu().numpy().view(np.uint16), cedge.cpu().numpy().view(np.uint16)
        for g in range(G):
            E = int(rowptr[g, n] & 0x7fff)
            rows, cols = meta[g, :E] >> 8, meta[g, :E] & 255
            assert colptr[g, 0] == 0 and colptr[g, n] == E
            for j in range(n):
                want = np.nonzero(cols == j)[0]                      # CSR order is row-major, so ascending ids = ascending rows
                got = cedge[g, colptr[g, j]:colptr[g, j + 1]]
                assert np.array_equal(got, want.astype(np.uint16)), (n, g, j)
                assert np.all(np.diff(rows[got].astype(np.int64)) > 0)


def test_precomputed_csr_is_bit_identical_to_in_kernel_csr():
    from tests.test_gpu_scoring import _setup
    cfg, sd, corpus, scorer = _setup(n_beh=50, seed=4)
    beh = torch.from_numpy(corpus.pair_behavior).cuda()
    news = torch.from_numpy(corpus.pair_news).cuda()
    out = {}
    for on in (True, False):
        scorer.enc.precompute_user_csr = on
        out[on] = (scorer.score_resident(beh, news), scorer.score_resident(beh, news, share_user_graphs=False))
    scorer.enc.precompute_user_csr = True
    assert torch.equal(out[True][0], out[False][0]) and torch.equal(out[True][1], out[False][1])
    assert torch.equal(out[True][0], out[True][1])
    m, b = _encoder('default_n3_L3')                                     # the module API on per-row adjacencies (no sharing)
    args = [b[k].repeat(64, *([1] * (b[k].dim() - 1))) for k in ORDER]
    res = {}
    for on in (True, False):
        m.precompute_user_csr = on
        res[on] = m.forward(*args)
    assert torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][1], res[False][1])
