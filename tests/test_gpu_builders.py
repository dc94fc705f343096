"""GPU parity of the integer builders and ranking kernels -- bit-exact against the oracle's
literal loops and against golden vectors produced by the unmodified reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import digat_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('H,C', [(50, 18), (7, 3), (33, 5)])     # (7,3),(33,5): (H+C)^2 not a multiple of 4 -> byte path
def test_user_graph_kernel_bit_exact(H, C):
    from digat_b200 import graphs
    rng = np.random.Generator(np.random.PCG64(11))
    N = 300
    lens = rng.integers(0, H + 1, size=N)
    lens[0], lens[1], lens[2] = 0, H, 1
    cats = rng.integers(0, C, size=(N, H))
    cats[3] = 0                                                    # a single category
    dev = torch.device('cuda:0')
    g, cm, ci = graphs.build_user_graphs_device(torch.from_numpy(cats).to(dev), torch.from_numpy(lens).to(dev), H, C)
    assert g.dtype == torch.bool and cm.dtype == torch.bool and ci.dtype == torch.int64
    g, cm, ci = g.cpu().numpy(), cm.cpu().numpy(), ci.cpu().numpy()
    for n in range(N):
        g0, cm0, ci0 = O.user_graph_loops(cats[n], int(lens[n]), H, C)
        assert np.array_equal(g[n], g0), n
        assert np.array_equal(cm[n], cm0) and np.array_equal(ci[n], ci0), n
    gv, cmv, civ = graphs.build_user_graphs(cats, lens, H, C)     # and the vectorised host builder
    assert np.array_equal(g, gv) and np.array_equal(cm, cmv) and np.array_equal(ci, civ)


def test_user_graph_kernel_rejects_bad_category():
    from digat_b200 import graphs
    dev = torch.device('cuda:0')
    cats = torch.zeros((2, 50), dtype=torch.int64, device=dev)
    cats[1, 0] = 18
    lens = torch.tensor([3, 3], device=dev)
    with pytest.raises(IndexError):
        graphs.build_user_graphs_device(cats, lens, 50, 18)
    with pytest.raises(IndexError):
        graphs.build_user_graphs_device(cats * 0, torch.tensor([3, 51], device=dev), 50, 18)


def test_sag_bfs_kernel_matches_reference_golden():
    from digat_b200 import graphs
    z = np.load(os.path.join(GOLDEN, 'sag_bfs.npz'))
    sim = z['sim']
    n_news = sim.shape[0]
    similar = [[(int(o), float(c)) for o, c in sim[k] if o >= 0] if k >= 1 else [] for k in range(n_news)]
    off, idx, cos = graphs.similar_to_csr(similar, n_news)
    node, graph, mask = graphs.sag_bfs_device(off, idx, cos, n_news, int(z['top_M']), int(z['hop']), int(z['n_nodes']),
                                              float(z['threshold']), device='cuda:0')
    assert node.dtype == torch.int32 and graph.dtype == torch.bool and mask.dtype == torch.bool
    assert np.array_equal(node.cpu().numpy(), z['node'])
    assert np.array_equal(graph.cpu().numpy(), z['graph'])
    assert np.array_equal(mask.cpu().numpy(), z['mask'])


@pytest.mark.parametrize('top_M,hop,n_news', [(3, 2, 500), (5, 2, 800), (8, 2, 700), (4, 3, 300), (3, 0, 50)])
def test_sag_bfs_kernel_matches_oracle_loops(top_M, hop, n_news):
    from digat_b200 import graphs, synth
    rng = np.random.Generator(np.random.PCG64(top_M * 100 + hop))
    n_nodes = synth.sag_size(top_M, hop) if hop > 0 else 1 + top_M
    if hop == 3:
        n_nodes = 1 + top_M + top_M * (top_M - 1) + top_M * (top_M - 1) ** 2
    thr = 0.5
    similar = [[]]
    for k in range(1, n_news):
        m = int(rng.integers(0, top_M + 1))
        others = rng.choice(np.arange(1, n_news), size=m, replace=False)
        cos = np.sort(rng.uniform(0.3, 1.0, size=m))[::-1]          # some below the threshold: cut at depth > 0
        similar.append([(int(o), float(c)) for o, c in zip(others, cos)])
    want = O.sag_bfs({k: similar[k] for k in range(1, n_news)}, n_news, top_M, hop, n_nodes, thr)
    off, idx, cos = graphs.similar_to_csr(similar, n_news)
    got = graphs.sag_bfs_device(off, idx, cos, n_news, top_M, hop, n_nodes, thr, device='cuda:0')
    for w, g in zip(want, got):
        assert np.array_equal(w, g.cpu().numpy())


def test_sag_bfs_kernel_flags_overflow():
    from digat_b200 import graphs
    similar = [[], [(2, 0.9), (3, 0.8), (4, 0.7)], [], [], []]
    off, idx, cos = graphs.similar_to_csr(similar, 5)
    with pytest.raises(IndexError):
        graphs.sag_bfs_device(off, idx, cos, 5, 3, 2, 3, 0.5, device='cuda:0')       # 4 nodes needed, 3 allowed


def test_rank_and_metrics_kernels_match_reference_golden():
    from digat_b200 import evaluate
    z = np.load(os.path.join(GOLDEN, 'metrics.npz'))
    dev = torch.device('cuda:0')
    off = torch.from_numpy(evaluate.impression_offsets(z['imp'])).to(dev)
    ranks = evaluate.rank_pairs_device(torch.from_numpy(z['scores']).to(dev), off)
    want = O.rank_lists(z['scores'], z['imp'])
    assert ranks.dtype == torch.int32
    assert ranks.cpu().tolist() == [r for lst in want for r in lst]                  # bit-exact, ties included
    m = evaluate.metrics_device(ranks, torch.from_numpy(z['labels']).to(dev), off)
    assert np.allclose(m, z['metrics'], atol=1e-12)                                  # reference tolerance is 1e-4


def test_rank_kernel_large_impressions_and_ties():
    from digat_b200 import evaluate
    rng = np.random.Generator(np.random.PCG64(5))
    sizes = np.concatenate([rng.integers(1, 400, size=200), [1, 2, 33, 64, 65, 1000]])
    imp = np.repeat(np.arange(len(sizes)), sizes)
    scores = rng.integers(-5, 6, size=imp.shape[0]).astype(np.float32) * 0.25       # heavy ties
    dev = torch.device('cuda:0')
    off = torch.from_numpy(evaluate.impression_offsets(imp)).to(dev)
    ranks = evaluate.rank_pairs_device(torch.from_numpy(scores).to(dev), off).cpu().numpy()
    want = evaluate.rank_lists(scores, imp)
    assert ranks.tolist() == [r for lst in want for r in lst]
    labels = (rng.random(imp.shape[0]) < 0.3).astype(np.int64)
    first = off.cpu().numpy()[:-1]
    labels[first] = 1
    big = sizes >= 2
    labels[first[big] + 1] = 0
    keep = np.repeat(big, sizes)                                                     # single-pair impressions: one class
    imp2 = np.cumsum(np.concatenate([[0], np.diff(imp[keep]) != 0]))
    off2 = torch.from_numpy(evaluate.impression_offsets(imp2)).to(dev)
    r2 = evaluate.rank_pairs_device(torch.from_numpy(scores[keep]).to(dev), off2)
    got = evaluate.metrics_device(r2, torch.from_numpy(labels[keep]).to(dev), off2)
    lab_lists = [labels[keep][imp2 == i].tolist() for i in range(int(imp2[-1]) + 1)]
    want_m = evaluate.metrics(evaluate.rank_lists(scores[keep], imp2), lab_lists)
    assert np.allclose(got, want_m, atol=1e-12)
    with pytest.raises(ValueError):                                                  # an impression with one class only
        evaluate.metrics_device(torch.from_numpy(ranks).to(dev), torch.ones(len(ranks), dtype=torch.int64, device=dev), off)


def test_compact_lists_matches_numpy():
    """DIGAT.compact_flags (digat_compact_lists under one event wait) against numpy nonzero / cumsum, with empty, full and
    missing lists."""
    from digat_b200.graphEncoders import DIGAT
    rng = np.random.Generator(np.random.PCG64(8))
    dev = torch.device('cuda:0')
    cases = [
        [rng.random(4096) < 0.03, rng.random((4096, 68)) < 0.5, None, rng.random((4096, 19)) < 0.4],
        [None, np.zeros((300, 10), dtype=bool), np.ones((7, 3), dtype=bool), None],
        [rng.random(1) < 2.0],
        [None, None],
    ]
    for flags in cases:
        got = DIGAT.compact_flags([None if f is None else torch.from_numpy(np.ascontiguousarray(f)).to(dev) for f in flags])
        assert len(got) == len(flags)
        for f, g in zip(flags, got):
            if f is None:
                assert g is None
                continue
            ids, pos = g
            want = np.flatnonzero(f.reshape(-1))
            assert ids.dtype == torch.int32 and pos.dtype == torch.int32
            assert np.array_equal(ids.cpu().numpy(), want)
            rank = np.cumsum(f.reshape(-1)) - 1
            sel = f.reshape(-1)
            assert np.array_equal(pos.cpu().numpy()[sel], rank[sel])
