"""Full-size run (BASELINE.json configs[1]: the MIND-small-dev-sized synthetic corpus of bench.py, 65 238 news, 73 152
behaviours, ~2.7 M pairs) checked through size-independent properties: the oracle cannot run at this size in seconds."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def full():
    import bench
    from digat_b200 import scoring
    from digat_b200.graphEncoders import DIGAT
    cfg, sd, corpus = bench.build_workload('mind_small_dev_n3_L3')
    enc = DIGAT(cfg, 400)
    enc.load_state_dict(sd)
    enc = enc.cuda().eval()
    scorer = scoring.Scorer(enc, corpus, 'cuda:0', build_user_graphs_on_device=True)
    scorer.cache_news_context()
    return corpus, scorer


def test_device_built_user_graphs_equal_host_builder_at_full_size(full):
    corpus, scorer = full
    assert scorer.user_graph.shape == corpus.user_graph.shape                   # [73152, 68, 68]
    for lo in range(0, corpus.user_graph.shape[0], 20000):                      # bit-exact, all 338 MB of it
        hi = min(lo + 20000, corpus.user_graph.shape[0])
        assert torch.equal(scorer.user_graph[lo:hi].cpu(), torch.from_numpy(corpus.user_graph[lo:hi]))
    assert torch.equal(scorer.cidx.cpu(), torch.from_numpy(corpus.user_category_indices))
    assert torch.equal(scorer.cmask.cpu(), torch.from_numpy(corpus.user_category_mask))


def test_full_corpus_scores_ranks_and_metrics(full):
    from digat_b200 import evaluate, scoring
    corpus, scorer = full
    n_pairs = corpus.pair_behavior.shape[0]
    out = scoring.evaluate_resident(scorer, corpus, batch_size=4096, on_single_class='skip')
    scores, ranks, off = out['scores'], out['ranks'], out['offsets']
    assert scores.shape == (n_pairs,) and bool(torch.isfinite(scores).all())
    # every impression's ranks are a permutation of 1..m: sum and sum of squares per impression (integer identities)
    m = (off[1:] - off[:-1]).to(torch.float64)
    seg = torch.repeat_interleave(torch.arange(off.shape[0] - 1, device=ranks.device), (off[1:] - off[:-1]))
    r = ranks.to(torch.float64)
    s1 = torch.zeros_like(m).index_add_(0, seg, r)
    s2 = torch.zeros_like(m).index_add_(0, seg, r * r)
    assert torch.equal(s1, m * (m + 1) / 2) and torch.equal(s2, m * (m + 1) * (2 * m + 1) / 6)
    # rank 1 holds the impression's maximum score
    mx = torch.full_like(m, -float('inf'), dtype=torch.float32).scatter_reduce_(0, seg, scores, 'amax')
    assert torch.equal(scores[ranks == 1], mx)
    assert all(0.0 <= v <= 1.0 for v in out['metrics'])
    # idempotence and batch-size independence of a slice (same rows -> same kernels -> identical bits for equal batches)
    beh = torch.from_numpy(corpus.pair_behavior[:8192]).cuda()
    news = torch.from_numpy(corpus.pair_news[:8192]).cuda()
    a = torch.cat([scorer.score_resident(beh[:4096], news[:4096]), scorer.score_resident(beh[4096:], news[4096:])])
    assert torch.equal(a, scores[:8192])
    # two shards of whole impressions reproduce the single-rank ranks
    parts = [scoring.evaluate_resident(scorer, corpus, batch_size=4096, rank=r_, world_size=2, on_single_class='skip')
             for r_ in range(2)]
    assert torch.equal(torch.cat([p['ranks'] for p in parts]), ranks)
    # host-batch path on a slice: same pairs through the DataLoader-shaped tensors (different batch composition for the
    # small context GEMMs only when batch sizes differ; here they are equal -> identical)
    hb = scoring.host_batch(corpus, np.arange(4096), pin=True)
    assert torch.equal(scorer.score_host_batch(*hb), scores[:4096])
