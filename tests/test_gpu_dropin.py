"""The reference's scoring hot loop, line for line (util.py:34-69: cached SAG gather, c_n0 cache over torch.narrow views in
chunks of batch_size news, a torch DataLoader yielding the 8-tuple of MIND_DevTest_Dataset, int64 news_ID, three
index_selects per batch, Model.inference), driven over digat_b200's Model / DIGAT exactly as the reference drives its own
-- and the same batches through Scorer.score_host_batch -- against the CPU oracle."""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader, Dataset

from oracle import digat_oracle as O
from tests.helpers import err_stats, fmt_stats, rel_err

pytestmark = pytest.mark.gpu
ORDER = ('news_graph_embeddings', 'news_graph', 'news_graph_mask', 'user_news_embedding', 'user_graph',
         'user_category_mask', 'user_category_indices')


class _DevTestDataset(Dataset):
    """The tuple reference MIND_dataset.py:97-105 returns per pair (user_graph_mask is produced and ignored by util.py:56)."""

    def __init__(self, corpus):
        self.c = corpus

    def __len__(self):
        return self.c.pair_behavior.shape[0]

    def __getitem__(self, i):
        b, nid = self.c.pair_behavior[i], self.c.pair_news[i]
        return (self.c.history[b].astype(np.int64), self.c.user_graph[b], np.zeros(1, dtype=bool), self.c.user_category_mask[b],
                self.c.user_category_indices[b], np.int64(nid), self.c.news_graph[nid], self.c.news_graph_mask[nid])


def _reference_style_compute_scores(model, corpus, batch_size):
    """util.compute_scores (util.py:34-69) with the news-encoder cache replaced by the synthetic embedding table."""
    model.eval()
    graph_encoder = model.graph_encoder
    max_history_num = graph_encoder.max_history_num
    cached_news_representations = torch.from_numpy(corpus.news_embeddings).cuda()
    cached_news_num, news_embedding_dim = cached_news_representations.shape
    with torch.no_grad():
        news_node_index = torch.from_numpy(corpus.news_node_ID).cuda()
        news_graph_masks = torch.from_numpy(corpus.news_graph_mask).cuda()
        cached_SA = cached_news_representations.index_select(dim=0, index=news_node_index.flatten()).view(
            [cached_news_num, -1, news_embedding_dim])
        cached_c_n0 = torch.zeros([cached_news_num, news_embedding_dim]).cuda()
        index = 0
        while index != cached_news_num:
            _index = min(index + batch_size, cached_news_num)
            batch_num = _index - index
            cached_c_n0[index:_index] = graph_encoder.compute_news_graph_context(
                torch.narrow(cached_SA, 0, index, batch_num), torch.narrow(news_graph_masks, 0, index, batch_num))
            index = _index
        dataloader = DataLoader(_DevTestDataset(corpus), batch_size=batch_size, shuffle=False, num_workers=0, pin_memory=True)
        scores = torch.zeros([len(corpus.pair_behavior)]).cuda()
        index = 0
        for (user_title_index, user_graph, _mask, user_category_mask, user_category_indices, news_ID, news_graph,
             news_graph_mask) in dataloader:
            user_title_index = user_title_index.cuda(non_blocking=True)
            user_graph = user_graph.cuda(non_blocking=True)
            user_category_mask = user_category_mask.cuda(non_blocking=True)
            user_category_indices = user_category_indices.cuda(non_blocking=True)
            news_ID = news_ID.cuda(non_blocking=True)
            news_graph = news_graph.cuda(non_blocking=True)
            news_graph_mask = news_graph_mask.cuda(non_blocking=True)
            bsz = user_title_index.size(0)
            user_representations = cached_news_representations.index_select(dim=0, index=user_title_index.flatten()).view(
                [bsz, max_history_num, news_embedding_dim])
            news_representations = cached_SA.index_select(dim=0, index=news_ID)
            c_n0 = cached_c_n0.index_select(dim=0, index=news_ID)
            scores[index:index + bsz] = model.inference(user_representations, user_graph, user_category_mask,
                                                        user_category_indices, news_representations, news_graph,
                                                        news_graph_mask, c_n0)
            index += bsz
    return scores


@pytest.mark.parametrize('batch_size', [64, 1024])
def test_reference_shaped_scoring_loop_matches_oracle(batch_size):
    from digat_b200 import scoring, synth
    from digat_b200.model import Model
    cfg = synth.make_config(SAG_neighbors=3, SAG_hops=2, graph_depth=3)
    sd = synth.make_state_dict(cfg, seed=21)
    corpus = synth.make_corpus(cfg, n_news=700, n_behaviors=60, mean_candidates=20.0, seed=8)
    model = Model(cfg, 400)
    model.graph_encoder.load_state_dict(sd)
    model = model.cuda()
    got = _reference_style_compute_scores(model, corpus, batch_size).cpu().numpy()
    P = O.cast_params(sd)
    n = corpus.pair_behavior.shape[0]
    ref, terms = [], []
    with torch.no_grad():
        for s in range(0, n, 64):
            b = synth.make_batch(corpus, np.arange(s, min(s + 64, n)))
            c0 = O.news_graph_context(P, b['news_graph_embeddings'], b['news_graph_mask'])
            cn, cu = O.inference(P, *[b[k] for k in ORDER], c0)
            ref.append(O.logits(cn, cu))
            terms.append((cn * cu).abs().sum(1))
    ref, terms = torch.cat(ref).numpy(), torch.cat(terms).numpy()
    st = err_stats(got, ref, terms)
    print('\nreference-shaped loop, batch %d, %d pairs: %s' % (batch_size, n, fmt_stats(st)))
    assert st['max_norm'] < 1e-5 and st['cond'] < 1e-5
    # the Scorer's host-batch entry point on the same DataLoader batches (int64 ids, pinned tensors)
    scorer = scoring.Scorer(model.graph_encoder, corpus, 'cuda:0')
    loader = DataLoader(_DevTestDataset(corpus), batch_size=batch_size, shuffle=False, num_workers=0, pin_memory=True)
    outs = []
    for (uti, ug, _m, cm, ci, nid, ng, nm) in loader:
        outs.append(scorer.score_host_batch(uti, ug, cm, ci, nid, ng, nm))
    scorer.check_index_errors()
    host = torch.cat(outs).cpu().numpy()
    assert rel_err(host, ref) < 1e-5
    # AUC / MRR / nDCG@5/10 from both score lists: equal to 1e-4 (README.md:64), through the reference's ranking rule
    labels = [corpus.labels[corpus.pair_behavior == i].tolist() for i in range(int(corpus.pair_behavior[-1]) + 1)]
    both = [i for i, lab in enumerate(labels) if len(set(lab)) == 2]           # sklearn's AUC needs both classes
    pick = lambda ranks: [ranks[i] for i in both]                             # noqa: E731
    m_got = O.metrics_from_ranks(pick(O.rank_lists(got, corpus.pair_behavior)), pick(labels))
    m_ref = O.metrics_from_ranks(pick(O.rank_lists(ref, corpus.pair_behavior)), pick(labels))
    assert max(abs(a - b) for a, b in zip(m_got, m_ref)) < 1e-4, (m_got, m_ref)
