"""Seeded synthetic MIND-shaped inputs and weights for the DIGAT encoder.

Everything is drawn from ``numpy.random.Generator(PCG64(seed))`` so the same tensors can be rebuilt on any box.
The invariants of the reference's data pipeline are kept:

* user graph / category mask / category indices follow the rule of reference MIND_corpus.py:146-176
  (identity diagonal, news-news edges inside a category, news-topic and topic-topic edges, padding bucket ``C``);
* SAG node table / adjacency / mask follow reference construct_SAG.py:449-485 + MIND_corpus.py:117-118,210
  (row 0 is the padding news, column 0 is the news itself, BFS-tree-like symmetric adjacency + identity,
  ``news_graph_mask[:, 0] = 0``);
* padded history slots point at news 0 (reference MIND_corpus.py:282).
"""
from dataclasses import dataclass
from types import SimpleNamespace

import math

import numpy as np
import torch

from . import graphs


def sag_size(neighbors: int, hops: int) -> int:
    """news_graph_size as derived in reference config.py:68-75."""
    size, fan = 1, 1
    for h in range(hops):
        fan *= neighbors if h == 0 else neighbors - 1
        size += fan
    return size


def make_config(SAG_neighbors=3, SAG_hops=2, graph_depth=3, max_history_num=50, category_num=18,
                dropout_rate=0.2, graph_encoder='DIGAT', news_encoder='MSA', **extra):
    """Duck-typed stand-in for reference config.Config (which needs a GPU and the datasets, config.py:84)."""
    return SimpleNamespace(SAG_neighbors=SAG_neighbors, SAG_hops=SAG_hops,
                           news_graph_size=sag_size(SAG_neighbors, SAG_hops), graph_depth=graph_depth,
                           max_history_num=max_history_num, category_num=category_num,
                           dropout_rate=dropout_rate, graph_encoder=graph_encoder, news_encoder=news_encoder, **extra)


def _xavier(rng, out_f, in_f, gain=1.0):
    bound = gain * np.sqrt(6.0 / (in_f + out_f))
    return rng.uniform(-bound, bound, size=(out_f, in_f)).astype(np.float32)


def make_state_dict(config, D=400, seed=0, trained_like=True):
    """state_dict with the parameter names/shapes of reference graphEncoders.py:21,52-73.

    ``trained_like`` makes biases and ``topic_node_embedding`` non-zero (the reference's ``initialize`` zeroes them,
    which would leave those code paths untested).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    relu_gain, leaky_gain = np.sqrt(2.0), np.sqrt(2.0 / (1 + 0.2 ** 2))
    sd = {}

    def bias(n):
        return (rng.normal(0, 0.05, size=n) if trained_like else np.zeros(n)).astype(np.float32)

    sd['topic_node_embedding'] = (rng.normal(0, 0.1, size=(config.category_num, D)) if trained_like
                                  else np.zeros((config.category_num, D))).astype(np.float32)
    for att in ('candidate_attention', 'userAttention'):
        sd[att + '.K.weight'] = _xavier(rng, D, D)
        sd[att + '.Q.weight'] = _xavier(rng, D, D)
        sd[att + '.Q.bias'] = bias(D)
    sd['news_graph_W.weight'] = _xavier(rng, D, 2 * D)
    sd['news_graph_W.bias'] = bias(D)
    sd['user_news_K.weight'] = _xavier(rng, D, D)
    sd['user_news_Q.weight'] = _xavier(rng, D, D)
    sd['user_news_Q.bias'] = bias(D)
    sd['featureAffine.weight'] = _xavier(rng, D, D, relu_gain)
    sd['featureAffine.bias'] = bias(D)
    for g in ('news', 'user'):
        for i in range(config.graph_depth):
            p = '%s_graph_attention_' % g
            sd[p + 'W.%d.weight' % i] = _xavier(rng, D, D)
            sd[p + 'W.%d.bias' % i] = bias(D)
            sd[p + 'ffn1.%d.weight' % i] = _xavier(rng, D, D, relu_gain)
            sd[p + 'ffn2.%d.weight' % i] = _xavier(rng, D, D, relu_gain)
            sd[p + 'ffn3.%d.weight' % i] = _xavier(rng, D, D, relu_gain)
            sd[p + 'ffn3.%d.bias' % i] = bias(D)
            sd[p + 'a.%d.weight' % i] = _xavier(rng, 1, D, leaky_gain)
    return {k: torch.from_numpy(v) for k, v in sd.items()}


# graph_encoder name -> (has news context, news layer kind, user layer kind); reference graphEncoders.py:201-842
ABLATIONS = {
    'wo_SA': (False, None, 'digat'),
    'Seq_SA': (True, None, 'digat'),
    'wo_interaction': (True, 'gat', 'gat'),
    'news_graph_wo_inter': (True, 'gat', 'digat'),
    'user_graph_wo_inter': (True, 'digat', 'gat'),
}


def make_ablation_state_dict(kind, config, D=400, seed=0):
    """Seeded trained-like state_dict of one ablation encoder (parameter names/shapes of reference
    graphEncoders.py:201-842; biases and topic nodes non-zero so that every code path matters)."""
    has_ctx, news_kind, user_kind = ABLATIONS[kind]
    rng = np.random.Generator(np.random.PCG64(seed + 77))
    relu_gain, leaky_gain = np.sqrt(2.0), np.sqrt(2.0 / (1 + 0.2 ** 2))
    bias = lambda n: rng.normal(0, 0.05, size=n).astype(np.float32)          # noqa: E731
    sd = {'topic_node_embedding': rng.normal(0, 0.1, size=(config.category_num, D)).astype(np.float32)}
    atts = (['candidate_attention'] if has_ctx else []) + ['userAttention']
    for att in atts:
        sd[att + '.K.weight'] = _xavier(rng, D, D)
        sd[att + '.Q.weight'] = _xavier(rng, D, D)
        sd[att + '.Q.bias'] = bias(D)
    if has_ctx:
        sd['news_graph_W.weight'] = _xavier(rng, D, 2 * D)
        sd['news_graph_W.bias'] = bias(D)
    sd['user_news_K.weight'] = _xavier(rng, D, D)
    sd['user_news_Q.weight'] = _xavier(rng, D, D)
    sd['user_news_Q.bias'] = bias(D)
    sd['featureAffine.weight'] = _xavier(rng, D, D, relu_gain)
    sd['featureAffine.bias'] = bias(D)
    for g, k in (('news', news_kind), ('user', user_kind)):
        p = '%s_graph_attention_' % g
        for i in range(config.graph_depth if k else 0):
            sd[p + 'W.%d.weight' % i] = _xavier(rng, D, D)
            sd[p + 'W.%d.bias' % i] = bias(D)
            if k == 'digat':
                sd[p + 'ffn1.%d.weight' % i] = _xavier(rng, D, D, relu_gain)
                sd[p + 'ffn2.%d.weight' % i] = _xavier(rng, D, D, relu_gain)
                sd[p + 'ffn3.%d.weight' % i] = _xavier(rng, D, D, relu_gain)
                sd[p + 'ffn3.%d.bias' % i] = bias(D)
                sd[p + 'a.%d.weight' % i] = _xavier(rng, 1, D, leaky_gain)
            else:
                sd[p + 'a1.%d.weight' % i] = _xavier(rng, 1, D, leaky_gain)
                sd[p + 'a2.%d.weight' % i] = _xavier(rng, 1, D, leaky_gain)
    return {k: torch.from_numpy(v) for k, v in sd.items()}


def make_text_config(vocabulary_size=500, max_title_length=32, word_embedding_dim=300, MSA_head_num=16, MSA_head_dim=25,
                     attention_dim=256, cnn_method='naive', cnn_kernel_num=400, cnn_window_size=3, **kw):
    """make_config plus the text-side fields of reference config.py (newsEncoders.MSA / CNN read them; config.py:43-45)."""
    return make_config(vocabulary_size=vocabulary_size, max_title_length=max_title_length,
                       word_embedding_dim=word_embedding_dim, MSA_head_num=MSA_head_num, MSA_head_dim=MSA_head_dim,
                       attention_dim=attention_dim, cnn_method=cnn_method, cnn_kernel_num=cnn_kernel_num,
                       cnn_window_size=cnn_window_size, word_threshold=3, dataset='synthetic', **kw)


def make_cnn_state_dict(config, seed=0):
    """Seeded trained-like state_dict of the CNN news encoder (names/shapes of reference newsEncoders.py:27-34, layers.py:7-26)."""
    rng = np.random.Generator(np.random.PCG64(seed + 523))
    E, Fk, A = config.word_embedding_dim, config.cnn_kernel_num, config.attention_dim
    sd = {'word_embedding.weight': rng.normal(0, 0.4, size=(config.vocabulary_size, E)).astype(np.float32)}
    sd['word_embedding.weight'][0] = 0                                   # padding token

    def conv(name, out_ch, window):
        bound = 1.0 / math.sqrt(E * window)
        sd['conv.%s.weight' % name] = rng.uniform(-bound, bound, size=(out_ch, E, window)).astype(np.float32)
        sd['conv.%s.bias' % name] = rng.uniform(-bound, bound, size=out_ch).astype(np.float32)
    if config.cnn_method == 'naive':
        conv('conv', Fk, config.cnn_window_size)
    elif config.cnn_method == 'group3':
        for name, window in (('conv1', 1), ('conv2', 3), ('conv3', 5)):
            conv(name, Fk // 3, window)
    else:
        raise ValueError('cnn_method %r' % config.cnn_method)
    sd['attention.affine1.weight'] = _xavier(rng, A, Fk, 5.0 / 3.0)
    sd['attention.affine1.bias'] = rng.normal(0, 0.05, size=A).astype(np.float32)
    sd['attention.affine2.weight'] = _xavier(rng, 1, A)
    return {k: torch.from_numpy(v) for k, v in sd.items()}


def make_msa_state_dict(config, seed=0):
    """Seeded trained-like state_dict of the MSA news encoder (names/shapes of reference newsEncoders.py:58-66)."""
    rng = np.random.Generator(np.random.PCG64(seed + 311))
    E, hd, A = config.word_embedding_dim, config.MSA_head_num * config.MSA_head_dim, config.attention_dim
    sd = {'word_embedding.weight': rng.normal(0, 0.4, size=(config.vocabulary_size, E)).astype(np.float32)}
    sd['word_embedding.weight'][0] = 0                                   # padding token
    for k in ('W_K', 'W_Q', 'W_V'):
        sd['multiheadSelfattention.%s.weight' % k] = _xavier(rng, hd, E)
    sd['multiheadSelfattention.W_Q.bias'] = rng.normal(0, 0.05, size=hd).astype(np.float32)
    sd['multiheadSelfattention.W_V.bias'] = rng.normal(0, 0.05, size=hd).astype(np.float32)
    sd['attention.affine1.weight'] = _xavier(rng, A, hd, 5.0 / 3.0)
    sd['attention.affine1.bias'] = rng.normal(0, 0.05, size=A).astype(np.float32)
    sd['attention.affine2.weight'] = _xavier(rng, 1, A)
    return {k: torch.from_numpy(v) for k, v in sd.items()}


def make_titles(config, n_titles, seed=0):
    """Token ids [n_titles, T] int64 (zero padded tail, a few empty titles) and their mask [n_titles, T] float (1 = token)."""
    rng = np.random.Generator(np.random.PCG64(seed + 911))
    T = config.max_title_length
    length = rng.integers(1, T + 1, size=n_titles)
    length[rng.random(n_titles) < 0.05] = 0                              # all-padding title (news 0 in the reference corpus)
    tok = rng.integers(1, config.vocabulary_size, size=(n_titles, T)).astype(np.int64)
    valid = np.arange(T)[None, :] < length[:, None]
    tok[~valid] = 0
    return torch.from_numpy(tok), torch.from_numpy(valid.astype(np.float32))


@dataclass
class Corpus:
    """Device-agnostic synthetic corpus: the arrays reference util.compute_scores reads from MIND_Corpus."""
    news_embeddings: np.ndarray          # [N_news, D] f32   (stands in for the cached news-encoder output, util.py:24-33)
    news_node_ID: np.ndarray             # [N_news, n_n] i32 (construct_SAG.py:452)
    news_graph: np.ndarray               # [N_news, n_n, n_n] bool
    news_graph_mask: np.ndarray          # [N_news, n_n] bool, [:,0]=0 (MIND_corpus.py:210)
    history: np.ndarray                  # [N_beh, H] i32 news ids, zero padded at the tail (MIND_corpus.py:282)
    history_category: np.ndarray         # [N_beh, H] i64 category per history slot, C for padding
    user_graph: np.ndarray               # [N_beh, n_u, n_u] bool
    user_category_mask: np.ndarray       # [N_beh, C+1] bool
    user_category_indices: np.ndarray    # [N_beh, H] i64
    pair_behavior: np.ndarray            # [N_pairs] i32 behaviour (impression) index of each pair, sorted
    pair_news: np.ndarray                # [N_pairs] i32 candidate news id of each pair
    labels: np.ndarray                   # [N_pairs] i8 synthetic click labels (>=1 positive per impression)


def make_sag(rng, n_news, n_n, neighbors, hops):
    """BFS-tree-shaped SAG tables with random truncation (news with few similar neighbours)."""
    node = np.zeros((n_news, n_n), dtype=np.int32)
    adj = np.zeros((n_news, n_n, n_n), dtype=bool)
    mask = np.zeros((n_news, n_n), dtype=bool)
    # parent position of each BFS slot for a full tree (fan-out `neighbors` at the root, neighbors-1 below)
    parent = [-1]
    frontier = [0]
    for h in range(hops):
        nxt = []
        for p in frontier:
            for _ in range(neighbors if h == 0 else neighbors - 1):
                parent.append(p)
                nxt.append(len(parent) - 1)
        frontier = nxt
    parent = np.array(parent[:n_n])
    present_len = rng.integers(1, n_n + 1, size=n_news)          # how many BFS slots are filled
    present_len[rng.random(n_news) < 0.05] = 1                    # isolated news: only itself
    ids = rng.integers(1, max(2, n_news), size=(n_news, n_n)).astype(np.int32)
    for pos in range(n_n):
        filled = pos < present_len
        node[:, pos] = np.where(filled, ids[:, pos], 0)
        mask[:, pos] = filled
        if pos > 0:
            rows = np.nonzero(filled)[0]
            adj[rows, parent[pos], pos] = True
            adj[rows, pos, parent[pos]] = True
    node[:, 0] = np.arange(n_news, dtype=np.int32)
    # a few cross edges between present nodes (construct_SAG.py:481-484 links already-visited nodes)
    extra = rng.integers(0, n_n, size=(n_news, 2))
    rows = np.nonzero((extra[:, 0] < present_len) & (extra[:, 1] < present_len))[0]
    adj[rows, extra[rows, 0], extra[rows, 1]] = True
    adj[rows, extra[rows, 1], extra[rows, 0]] = True
    adj |= np.eye(n_n, dtype=bool)[None]                          # MIND_corpus.py:117-118
    node[0] = 0; adj[0] = np.eye(n_n, dtype=bool); mask[0] = False  # padding news (construct_SAG.py:456 starts at 1)
    mask[:, 0] = False                                            # MIND_corpus.py:210
    return node, adj, mask


def make_corpus(config, D=400, n_news=2000, n_behaviors=64, mean_candidates=8.0, seed=0, emb_scale=0.3,
                nonneg=False, behavior_seed=None, build_user_graph=True) -> Corpus:
    """behavior_seed: draw the behaviours / pairs from their own generator (ranks of a sharded run then hold the SAME news
    side and DIFFERENT behaviour shards); None keeps the single generator sequence the golden fixtures were made with.
    build_user_graph=False leaves user_graph / user_category_mask / user_category_indices as None (at MIND-large size the
    [N_beh, 68, 68] array is 11 GB on the host): build them on the device (Scorer(build_user_graphs_on_device=True)) or per
    batch (user_graphs_of)."""
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    H, C, n_n = config.max_history_num, config.category_num, config.news_graph_size
    emb = rng.normal(0, emb_scale, size=(n_news, D)).astype(np.float32)
    if nonneg:
        emb = np.maximum(emb, 0)
    news_cat = _zipf_categories(rng, n_news, C)
    node, adj, mask = make_sag(rng, n_news, n_n, config.SAG_neighbors, config.SAG_hops)
    if behavior_seed is not None:
        rng = np.random.Generator(np.random.PCG64(behavior_seed + 5000))
    hist_len = rng.integers(0, H + 1, size=n_behaviors)
    hist_len[rng.random(n_behaviors) < 0.03] = 0
    history = rng.integers(1, n_news, size=(n_behaviors, H)).astype(np.int32)
    history[np.arange(H)[None, :] >= hist_len[:, None]] = 0
    hist_cat = np.where(np.arange(H)[None, :] < hist_len[:, None], news_cat[history], C).astype(np.int64)
    ug, cmask, cidx = graphs.build_user_graphs(hist_cat, hist_len, H, C) if build_user_graph else (None, None, None)
    n_cand = np.maximum(1, rng.geometric(1.0 / mean_candidates, size=n_behaviors))
    pair_beh = np.repeat(np.arange(n_behaviors, dtype=np.int32), n_cand)
    pair_news = rng.integers(1, n_news, size=pair_beh.shape[0]).astype(np.int32)
    labels = (rng.random(pair_beh.shape[0]) < 0.1).astype(np.int8)
    first = np.concatenate([[0], np.cumsum(n_cand)[:-1]])
    labels[first] = 1                                            # every impression has a positive ...
    two = n_cand >= 2
    labels[first[two] + 1] = 0                                   # ... and, when it can, a negative (AUC defined)
    return Corpus(emb, node, adj, mask, history, hist_cat, ug, cmask, cidx, pair_beh, pair_news, labels)


def user_graphs_of(corpus: Corpus, config, behaviors: np.ndarray):
    """(user_graph, category_mask, category_indices) of the given behaviours: slices of the corpus arrays, or built on the
    fly from the history categories when the corpus was made with build_user_graph=False."""
    if corpus.user_graph is not None:
        return corpus.user_graph[behaviors], corpus.user_category_mask[behaviors], corpus.user_category_indices[behaviors]
    H, C = config.max_history_num, config.category_num
    cat = corpus.history_category[behaviors]
    return graphs.build_user_graphs(cat, (cat < C).sum(axis=1), H, C)


def _zipf_categories(rng, n, C):
    w = 1.0 / np.arange(1, C + 1)
    return rng.choice(C, size=n, p=w / w.sum()).astype(np.int64)


def make_batch(corpus: Corpus, pair_ids: np.ndarray):
    """Host-side tensors of one scoring batch, exactly what reference util.py:56-68 hands to Model.inference
    (after its three index_selects)."""
    b = corpus.pair_behavior[pair_ids]
    nid = corpus.pair_news[pair_ids]
    t = torch.from_numpy
    emb = corpus.news_embeddings
    return dict(
        user_news_embedding=t(emb[corpus.history[b]]),                       # [B,H,D]
        user_graph=t(corpus.user_graph[b]),                                  # [B,n_u,n_u] bool
        user_category_mask=t(corpus.user_category_mask[b]),                  # [B,C+1] bool
        user_category_indices=t(corpus.user_category_indices[b]),            # [B,H] i64
        news_graph_embeddings=t(emb[corpus.news_node_ID[nid]]),              # [B,n_n,D]
        news_graph=t(corpus.news_graph[nid]),                                # [B,n_n,n_n] bool
        news_graph_mask=t(corpus.news_graph_mask[nid]),                      # [B,n_n] bool
    )
