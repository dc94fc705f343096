"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the DIGAT dual-graph encoder (the parity oracle).

A functional re-statement, in plain torch CPU ops, of the arithmetic of the reference hot path.  It is the checker
for the CUDA path and the ``cpu_baseline`` ("port") timed by bench.py; it is never imported by ``digat_b200/``.

Each function cites the reference lines it follows (paths relative to /root/reference):

* ``sdpa``               -- layers.py:199-206            (ScaledDotProductAttention.forward)
* ``news_graph_context`` -- graphEncoders.py:109-114
* ``user_graph_context`` -- graphEncoders.py:123-134     (torch_scatter ops restated in oracle/scatter_shim.py)
* ``graph_layer``        -- graphEncoders.py:143-154 / 163-174   (Eq. (8) at :150 / :170)
* ``inference``          -- graphEncoders.py:189-198  + model.py:87-90 (logits)
* ``forward``            -- graphEncoders.py:177-187  + model.py:73-77 (eval / p=0 semantics, dropout omitted)
* ``gat_layer`` / ``ablation_inference`` / ``ablation_forward`` -- the five ablation encoders, graphEncoders.py:201-842
* ``msa_news_encoder``   -- newsEncoders.py:58-82 + layers.py:50-115 (MSA title encoder, eval mode)
* ``gather_*``           -- util.py:34-36, 65-67
* ``rank_lists`` / ``metrics`` -- util.py:70-80 + evaluate.py:32-89
* ``user_graph_loops``   -- MIND_corpus.py:143-176 (literal loops; integer oracle)
* ``sag_bfs``            -- construct_SAG.py:449-485 (literal BFS; integer oracle)

Pinning: the reference ships no tests, fixtures or golden vectors for this path.  The
restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF: ``oracle/make_golden.py`` imports the unmodified
/root/reference modules (torch_scatter replaced by the shim, whose algorithm is third-party and therefore
"parity unpinned" at that one boundary) and commits their outputs under tests/golden/; tests/test_oracle.py
checks this file against those vectors bit-for-bit.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .scatter_shim import scatter_softmax, scatter_sum

NEG_FILL = -1e9          # masked_fill value, graphEncoders.py:152 / layers.py:202 (finite on purpose)
LEAKY_SLOPE = 0.2        # graphEncoders.py:20


def cast_params(state_dict, dtype=torch.float32):
    return {k: v.to(dtype) for k, v in state_dict.items()}


def _lin(x, w, b=None):
    return F.linear(x, w, b)


def sdpa(P, prefix, feature, query, mask):
    D = feature.shape[-1]
    keys = _lin(feature, P[prefix + '.K.weight'])
    q = _lin(query, P[prefix + '.Q.weight'], P[prefix + '.Q.bias']).unsqueeze(2)
    a = torch.bmm(keys, q).squeeze(2) / math.sqrt(float(D))
    alpha = F.softmax(a.masked_fill(mask == 0, NEG_FILL), dim=1)
    return torch.bmm(alpha.unsqueeze(1), feature).squeeze(1)


def news_graph_context(P, X, mask):
    local = X.select(1, 0)
    glob = sdpa(P, 'candidate_attention', X, local, mask)
    gate = torch.sigmoid(_lin(torch.cat([local, glob], dim=1), P['news_graph_W.weight'], P['news_graph_W.bias']))
    return gate * local + (1 - gate) * glob


def user_graph_context(P, X_u, cat_mask, cat_idx, c_n, H):
    Xh = X_u[:, :H, :]
    D = Xh.shape[-1]
    n_bucket = cat_mask.shape[1]
    K = _lin(Xh, P['user_news_K.weight'])
    Q = _lin(c_n, P['user_news_Q.weight'], P['user_news_Q.bias']).unsqueeze(2)
    a = torch.bmm(K, Q).squeeze(2) / math.sqrt(float(D))
    alpha = scatter_softmax(a, cat_idx, 1, dim_size=n_bucket).unsqueeze(2)
    T = scatter_sum(alpha * Xh, cat_idx, dim=1, dim_size=n_bucket)
    T = F.relu(_lin(T, P['featureAffine.weight'], P['featureAffine.bias'])) + T
    return sdpa(P, 'userAttention', T, c_n, cat_mask)


def attention_scores(P, g, i, X, ctx):
    """Eq. (8) logits s_ij (before leaky-relu); the [B,n,n,D] tensor IS materialised here, as in the reference."""
    p = '%s_graph_attention_' % g
    B, n, D = X.shape
    K1 = _lin(X, P[p + 'ffn1.%d.weight' % i]).unsqueeze(1)
    K2 = _lin(X, P[p + 'ffn2.%d.weight' % i]).unsqueeze(2)
    K3 = _lin(ctx, P[p + 'ffn3.%d.weight' % i], P[p + 'ffn3.%d.bias' % i]).view(B, 1, 1, D)
    return _lin(F.relu(K3 + K1 + K2), P[p + 'a.%d.weight' % i]).squeeze(3)


def graph_layer(P, g, i, X, adj, ctx):
    p = '%s_graph_attention_' % g
    h = _lin(X, P[p + 'W.%d.weight' % i], P[p + 'W.%d.bias' % i])
    e = F.leaky_relu(attention_scores(P, g, i, X, ctx), LEAKY_SLOPE)
    alpha = F.softmax(e.masked_fill(adj == 0, NEG_FILL), dim=2)
    return F.relu(torch.bmm(alpha, h)) + X


def _user_nodes(P, user_news_embedding):
    B = user_news_embedding.shape[0]
    topic = P['topic_node_embedding'].unsqueeze(0).expand(B, -1, -1)
    return torch.cat([user_news_embedding, topic], dim=1)


def _layers(P, L, H, X_n, adj_n, mask_n, X_u, adj_u, cat_mask, cat_idx, c_n, c_u):
    for i in range(L):
        X_n_new = graph_layer(P, 'news', i, X_n, adj_n, c_u)
        X_u = graph_layer(P, 'user', i, X_u, adj_u, c_n)
        X_n = X_n_new
        c_n = c_n + news_graph_context(P, X_n, mask_n)
        c_u = c_u + user_graph_context(P, X_u, cat_mask, cat_idx, c_n, H)
    return c_n, c_u


def depth_of(P):
    return 1 + max(int(k.split('.')[1]) for k in P if k.startswith('news_graph_attention_W.'))


def inference(P, news_graph_embeddings, news_graph, news_graph_mask, user_news_embedding, user_graph,
              user_category_mask, user_category_indices, news_graph_context_0):
    """-> (news_ctx [B,D], user_ctx [B,D]); same argument order as reference DIGAT.inference."""
    H = user_news_embedding.shape[1]
    X_u = _user_nodes(P, user_news_embedding)
    c_u = user_graph_context(P, X_u, user_category_mask, user_category_indices, news_graph_context_0, H)
    return _layers(P, depth_of(P), H, news_graph_embeddings, news_graph, news_graph_mask, X_u, user_graph,
                   user_category_mask, user_category_indices, news_graph_context_0, c_u)


def forward(P, news_graph_embeddings, news_graph, news_graph_mask, user_news_embedding, user_graph,
            user_category_mask, user_category_indices):
    """Reference DIGAT.forward with every dropout an identity (eval mode, or train mode with p=0)."""
    c_n0 = news_graph_context(P, news_graph_embeddings, news_graph_mask)
    return inference(P, news_graph_embeddings, news_graph, news_graph_mask, user_news_embedding, user_graph,
                     user_category_mask, user_category_indices, c_n0)


def logits(news_ctx, user_ctx):
    return (user_ctx * news_ctx).sum(dim=1)


# ----------------------------------------------------------------------------- ablation encoders (graphEncoders.py:201-842)
def gat_layer(P, g, i, X, adj):
    """Vanilla-GAT layer of wo_interaction / News_graph_wo_inter / User_graph_wo_inter (graphEncoders.py:494-503, 511-520,
    640-649, 807-816): a1 broadcasts over rows (neighbour term), a2 over columns (query term)."""
    p = '%s_graph_attention_' % g
    B, n, _ = X.shape
    h = _lin(X, P[p + 'W.%d.weight' % i], P[p + 'W.%d.bias' % i])
    a1 = _lin(h, P[p + 'a1.%d.weight' % i]).view(B, 1, n)
    a2 = _lin(h, P[p + 'a2.%d.weight' % i])
    e = F.leaky_relu(a1 + a2, LEAKY_SLOPE)
    alpha = F.softmax(e.masked_fill(adj == 0, NEG_FILL), dim=2)
    return F.relu(torch.bmm(alpha, h)) + X


ABLATION_LAYERS = {            # name -> (news layer kind, user layer kind)
    'wo_interaction': ('gat', 'gat'), 'news_graph_wo_inter': ('gat', 'digat'), 'user_graph_wo_inter': ('digat', 'gat'),
}


def ablation_inference(kind, P, news_graph_embeddings, news_graph, news_graph_mask, user_news_embedding, user_graph,
                       user_category_mask, user_category_indices, news_graph_context_0):
    """``inference`` of the five ablation encoders; same argument order as the reference."""
    H = user_news_embedding.shape[1]
    X_n, X_u = news_graph_embeddings, _user_nodes(P, user_news_embedding)
    L = 1 + max(int(k.split('.')[1]) for k in P if k.startswith('user_graph_attention_W.'))
    if kind == 'wo_SA':                                                       # graphEncoders.py:288-295
        cand = X_n.select(1, 0)
        for i in range(L):
            X_u = graph_layer(P, 'user', i, X_u, user_graph, cand)
        return cand, user_graph_context(P, X_u, user_category_mask, user_category_indices, cand, H)
    c_n = news_graph_context_0
    c_u = user_graph_context(P, X_u, user_category_mask, user_category_indices, c_n, H)
    if kind == 'Seq_SA':                                                      # graphEncoders.py:400-407
        for i in range(L):
            X_u = graph_layer(P, 'user', i, X_u, user_graph, c_n)
            c_u = c_u + user_graph_context(P, X_u, user_category_mask, user_category_indices, c_n, H)
        return c_n, c_u
    news_kind, user_kind = ABLATION_LAYERS[kind]                              # graphEncoders.py:537-548, 685-695, 832-842
    for i in range(L):
        X_n_new = gat_layer(P, 'news', i, X_n, news_graph) if news_kind == 'gat' else graph_layer(P, 'news', i, X_n, news_graph, c_u)
        X_u = gat_layer(P, 'user', i, X_u, user_graph) if user_kind == 'gat' else graph_layer(P, 'user', i, X_u, user_graph, c_n)
        X_n = X_n_new
        c_n = c_n + news_graph_context(P, X_n, news_graph_mask)
        c_u = c_u + user_graph_context(P, X_u, user_category_mask, user_category_indices, c_n, H)
    return c_n, c_u


def ablation_forward(kind, P, news_graph_embeddings, news_graph, news_graph_mask, *user_args):
    """``forward`` (eval / p=0): the initial news context is computed instead of read from the cache."""
    c_n0 = None if kind == 'wo_SA' else news_graph_context(P, news_graph_embeddings, news_graph_mask)
    return ablation_inference(kind, P, news_graph_embeddings, news_graph, news_graph_mask, *user_args, c_n0)


# ----------------------------------------------------------------------------- MSA news encoder (newsEncoders.py:58-82)
def msa_news_encoder(P, title_text, title_mask, heads, dk):
    """title_text [B, news_num, T] int64, title_mask [B, news_num, T] -> [B, news_num, heads*dk]; eval mode (no dropout).
    Multi-head self-attention without padding mask (layers.py:78-97), relu, additive attention pooling (layers.py:107-115)."""
    B, news_num, T = title_text.shape
    n = B * news_num
    mask = title_mask.view(n, T)
    w = F.embedding(title_text, P['word_embedding.weight']).view(n, T, -1)
    pre = 'multiheadSelfattention.'
    Q = _lin(w, P[pre + 'W_Q.weight'], P[pre + 'W_Q.bias']).view(n, T, heads, dk).transpose(1, 2).contiguous().view(n * heads, T, dk)
    K = _lin(w, P[pre + 'W_K.weight']).view(n, T, heads, dk).transpose(1, 2).contiguous().view(n * heads, T, dk)
    V = _lin(w, P[pre + 'W_V.weight'], P[pre + 'W_V.bias']).view(n, T, heads, dk).transpose(1, 2).contiguous().view(n * heads, T, dk)
    A = torch.bmm(Q, K.transpose(1, 2).contiguous()) / math.sqrt(float(dk))
    out = torch.bmm(F.softmax(A, dim=2), V).view(n, heads, T, dk).transpose(1, 2).contiguous().view(n, T, heads * dk)
    h = F.relu(out)
    att = torch.tanh(_lin(h, P['attention.affine1.weight'], P['attention.affine1.bias']))
    a = _lin(att, P['attention.affine2.weight']).squeeze(dim=2)
    alpha = F.softmax(a.masked_fill(mask == 0, NEG_FILL), dim=1).unsqueeze(dim=1)
    return torch.bmm(alpha, h).squeeze(dim=1).view(B, news_num, heads * dk)


def cnn_news_encoder(P, title_text, title_mask, method='naive'):
    """CNN news encoder, eval mode (newsEncoders.py:41-54): word embeddings, Conv1D + relu (layers.py:36-41; 'naive' = one
    nn.Conv1d with padding (window-1)//2, 'group3' = windows 1/3/5 concatenated over the channels), additive attention pooling
    (layers.py:107-115).  title_text [B, news_num, T] int64 -> [B, news_num, cnn_kernel_num]."""
    B, news_num, T = title_text.shape
    n = B * news_num
    mask = title_mask.view(n, T)
    w = F.embedding(title_text, P['word_embedding.weight']).view(n, T, -1).permute(0, 2, 1)

    def conv(name, pad):
        return F.conv1d(w, P['conv.%s.weight' % name], P['conv.%s.bias' % name], padding=pad)
    if method == 'naive':
        c = conv('conv', (P['conv.conv.weight'].shape[2] - 1) // 2)
    elif method == 'group3':
        c = torch.cat([conv('conv1', 0), conv('conv2', 1), conv('conv3', 2)], dim=1)
    else:
        raise ValueError(method)
    h = F.relu(c).permute(0, 2, 1)
    att = torch.tanh(_lin(h, P['attention.affine1.weight'], P['attention.affine1.bias']))
    a = _lin(att, P['attention.affine2.weight']).squeeze(dim=2)
    alpha = F.softmax(a.masked_fill(mask == 0, NEG_FILL), dim=1).unsqueeze(dim=1)
    return torch.bmm(alpha, h).squeeze(dim=1).view(B, news_num, -1)


# ----------------------------------------------------------------------------- gathers (util.py:34-36, 65-67)
def gather_sag_nodes(news_table, news_node_ID):
    n_news, n_n = news_node_ID.shape
    return news_table.index_select(0, news_node_ID.flatten().long()).view(n_news, n_n, -1)


def gather_rows(table, index):
    return table.index_select(0, index.flatten().long()).view(*index.shape, -1)


# ----------------------------------------------------------------------------- ranking + metrics
def rank_lists(scores, impression_of_pair):
    """Per-impression rank lists exactly as reference util.py:70-80 writes them (stable descending sort)."""
    scores = [float(s) for s in scores]
    n_imp = int(impression_of_pair[-1]) + 1
    sub = [[] for _ in range(n_imp)]
    for i, imp in enumerate(impression_of_pair):
        sub[int(imp)].append([scores[i], len(sub[int(imp)])])
    out = []
    for s in sub:
        s.sort(key=lambda x: x[0], reverse=True)
        r = [0] * len(s)
        for j in range(len(s)):
            r[s[j][1]] = j + 1
        out.append(r)
    return out


def metrics_from_ranks(ranks, labels_per_impression):
    """AUC / MRR / nDCG@5 / nDCG@10 following reference evaluate.py:7-30,66-89 (score = 1/rank)."""
    from sklearn.metrics import roc_auc_score

    def dcg(y_true, y_score, k):
        order = np.argsort(y_score)[::-1]
        y = np.take(y_true, order[:k])
        return np.sum((2 ** y - 1) / np.log2(np.arange(len(y)) + 2))

    aucs, mrrs, n5, n10 = [], [], [], []
    for r, lab in zip(ranks, labels_per_impression):
        if len(lab) == 0:
            continue
        y_true = np.array(lab, dtype='float32')
        y_score = [1.0 / x for x in r]
        aucs.append(roc_auc_score(y_true, y_score))
        order = np.argsort(y_score)[::-1]
        yt = np.take(y_true, order)
        mrrs.append(np.sum(yt / (np.arange(len(yt)) + 1)) / np.sum(yt))
        n5.append(dcg(y_true, y_score, 5) / dcg(y_true, y_true, 5))
        n10.append(dcg(y_true, y_score, 10) / dcg(y_true, y_true, 10))
    return float(np.mean(aucs)), float(np.mean(mrrs)), float(np.mean(n5)), float(np.mean(n10))


# ----------------------------------------------------------------------------- integer oracles (literal loops)
def user_graph_loops(history_category, history_len, H, C):
    """Literal restatement of MIND_corpus.py:143-176 for ONE behaviour.
    history_category: the categories of the (already truncated to the last H) clicked news, in order."""
    n = H + C
    g = np.identity(n, dtype=bool)
    cmask = np.zeros(C + 1, dtype=bool)
    cidx = np.full([H], C, dtype=np.int64)
    for i in range(history_len):
        ci = int(history_category[i])
        cmask[ci] = 1
        cidx[i] = ci
        g[i, H + ci] = 1
        g[H + ci, i] = 1
        for j in range(i + 1, history_len):
            cj = int(history_category[j])
            if ci == cj:
                g[i, j] = 1
                g[j, i] = 1
            else:
                g[H + ci, H + cj] = 1
                g[H + cj, H + ci] = 1
    return g, cmask, cidx


def sag_bfs(similar, n_news, top_M, hop, n_nodes, threshold):
    """Literal restatement of construct_SAG.py:449-485 on integer ids.
    ``similar[k]`` = list of (news_index, cos_similarity) sorted by similarity, for news k (k >= 1)."""
    node = np.zeros([n_news, n_nodes], dtype=np.int32)
    graph = np.zeros([n_news, n_nodes, n_nodes], dtype=bool)
    mask = np.zeros([n_news, n_nodes], dtype=bool)
    mask[:, 0] = 1
    for i in range(1, n_news):
        node[i, 0] = i
        pos_of = {i: 0}
        depth = [0] * n_nodes
        head, rear = 0, 1
        while head < rear:
            if depth[head] == hop:
                head += 1
                continue
            cur = int(node[i, head])
            for k, (other, cos) in enumerate(similar[cur]):
                if depth[head] > 0 and (cos < threshold or k == top_M - 1):
                    break
                if other not in pos_of:
                    node[i, rear] = other
                    mask[i, rear] = 1
                    pos_of[other] = rear
                    graph[i, head, rear] = 1
                    graph[i, rear, head] = 1
                    depth[rear] = depth[head] + 1
                    rear += 1
                else:
                    p = pos_of[other]
                    graph[i, head, p] = 1
                    graph[i, p, head] = 1
            head += 1
    return node, graph, mask
