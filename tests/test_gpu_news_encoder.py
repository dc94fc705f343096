"""GPU parity of the MSA news (title) encoder (reference newsEncoders.py:58-82) against the
golden vectors of the unmodified reference class, small (exact-fp32 GEMMs) and replicated (tcgen05 GEMMs, K = 300 with a
partial last k-block), plus ragged cases against the CPU oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import digat_oracle as O
from tests.helpers import GOLDEN, msa_inputs, rel_err, sha

pytestmark = pytest.mark.gpu


def _encoder(cfg, sd):
    from digat_b200.newsEncoders import MSA
    m = MSA(cfg)
    m.load_state_dict(sd)
    return m.cuda().eval()


@pytest.mark.parametrize('rep', [1, 40])
def test_msa_matches_reference_golden(rep):
    cfg, sd, tok, mask = msa_inputs()
    z = np.load(os.path.join(GOLDEN, 'news_encoder_msa.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    assert meta['x:title_text'] == sha(tok.numpy()) and all(meta['w:' + k] == sha(v.numpy()) for k, v in sd.items())
    m = _encoder(cfg, sd)
    with torch.no_grad():
        out = m(tok.cuda().repeat(rep, 1, 1), mask.cuda().repeat(rep, 1, 1))
    torch.cuda.synchronize()
    assert out.shape == (4 * rep, 6, 400)
    got = out.cpu().numpy()
    for r in (0, rep - 1):
        e32 = rel_err(got[4 * r:4 * r + 4], z['ref32_news'])
        assert e32 < 1e-5, 'x%d replica %d: rel err vs fp32 reference %.3e (vs fp64 %.3e)' % (
            rep, r, e32, rel_err(got[4 * r:4 * r + 4], z['ref64_news']))


@pytest.mark.parametrize('T,heads,dk,E,A,n', [(32, 16, 25, 300, 256, 700), (20, 4, 16, 100, 64, 33), (7, 2, 32, 52, 200, 5)])
def test_msa_matches_oracle_other_shapes(T, heads, dk, E, A, n):
    from digat_b200 import synth
    cfg = synth.make_text_config(vocabulary_size=300, max_title_length=T, word_embedding_dim=E, MSA_head_num=heads,
                                 MSA_head_dim=dk, attention_dim=A)
    sd = synth.make_msa_state_dict(cfg, seed=T)
    tok, mask = synth.make_titles(cfg, n, seed=T + 1)
    m = _encoder(cfg, sd)
    with torch.no_grad():
        got = m(tok.cuda().view(1, n, T), mask.cuda().view(1, n, T)).cpu()
        ref = O.msa_news_encoder(O.cast_params(sd, torch.float64), tok.view(1, n, T), mask.double().view(1, n, T), heads, dk)
    assert rel_err(got.numpy(), ref.numpy()) < 1e-5


def test_msa_feeds_the_graph_encoder_through_model_forward():
    """reference Model.forward on token tensors (model.py:54-77): news encoder -> DIGAT -> logits, against the oracle."""
    from digat_b200 import synth
    from digat_b200.model import Model
    cfg = synth.make_text_config(graph_depth=2)
    model = Model(cfg)
    assert model.news_encoder is not None and model.news_embedding_dim == 400
    sd_g = synth.make_state_dict(cfg, seed=3)
    sd_n = synth.make_msa_state_dict(cfg, seed=4)
    model.graph_encoder.load_state_dict(sd_g)
    model.news_encoder.load_state_dict(sd_n)
    model = model.cuda().eval()
    corpus = synth.make_corpus(cfg, n_news=60, n_behaviors=6, mean_candidates=3.0, seed=2)
    bs, news_num, n_n, H, T = 3, 2, cfg.news_graph_size, 50, cfg.max_title_length
    tok, mask = synth.make_titles(cfg, 60, seed=9)
    beh = np.arange(bs)
    cand = np.array([[5, 9], [11, 3], [7, 20]])
    u_tok, u_mask = tok[corpus.history[beh]], mask[corpus.history[beh]]
    n_tok, n_mask = tok[corpus.news_node_ID[cand]], mask[corpus.news_node_ID[cand]]
    t = torch.from_numpy
    ug, cm, ci = t(corpus.user_graph[beh]), t(corpus.user_category_mask[beh]), t(corpus.user_category_indices[beh])
    ng, nm = t(corpus.news_graph[cand]), t(corpus.news_graph_mask[cand])
    with torch.no_grad():
        got = model(u_tok.cuda(), u_mask.cuda(), ug.cuda(), cm.cuda(), ci.cuda(), n_tok.cuda(), n_mask.cuda(), ng.cuda(),
                    nm.cuda()).cpu()
        Pn, Pg = O.cast_params(sd_n), O.cast_params(sd_g)
        c_emb = O.msa_news_encoder(Pn, n_tok.view(bs * news_num, n_n, T), n_mask.view(bs * news_num, n_n, T), 16, 25)
        u_emb = O.msa_news_encoder(Pn, u_tok, u_mask, 16, 25)
        ex = lambda x: x.unsqueeze(1).expand(-1, news_num, *([-1] * (x.dim() - 1))).reshape(bs * news_num, *x.shape[1:])  # noqa: E731
        cn, cu = O.forward(Pg, c_emb, ng.view(bs * news_num, n_n, n_n), nm.view(bs * news_num, n_n), ex(u_emb), ex(ug), ex(cm), ex(ci))
        ref = O.logits(cn, cu).view(bs, news_num)
    assert got.shape == (bs, news_num)
    assert rel_err(got.numpy(), ref.numpy()) < 1e-5


@pytest.mark.parametrize('method,rep', [('naive', 1), ('naive', 30), ('group3', 1)])
def test_cnn_matches_reference_golden(method, rep):
    """CNN news encoder (reference newsEncoders.py:27-55): im2col gather + one GEMM with relu + additive pooling against the
    unmodified reference class, cnn_method naive (window 3) and group3 (windows 1/3/5 as one packed window-5 product)."""
    from digat_b200.newsEncoders import CNN
    from tests.helpers import cnn_inputs
    cfg, sd, tok, mask = cnn_inputs(method)
    z = np.load(os.path.join(GOLDEN, 'news_encoder_cnn_%s.npz' % method))
    meta = json.loads(bytes(z['meta']).decode())
    assert meta['x:title_text'] == sha(tok.numpy()) and all(meta['w:' + k] == sha(v.numpy()) for k, v in sd.items())
    m = CNN(cfg)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    with torch.no_grad():
        out = m(tok.cuda().repeat(rep, 1, 1), mask.cuda().repeat(rep, 1, 1))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert got.shape == (4 * rep, 6, cfg.cnn_kernel_num)
    for r in (0, rep - 1):
        e32 = rel_err(got[4 * r:4 * r + 4], z['ref32_news'])
        assert e32 < 1e-5, 'x%d replica %d: rel err vs fp32 reference %.3e (vs fp64 %.3e)' % (
            rep, r, e32, rel_err(got[4 * r:4 * r + 4], z['ref64_news']))


@pytest.mark.parametrize('method,T,E,Fk,A,window,n', [('naive', 20, 100, 64, 64, 5, 33), ('naive', 7, 52, 48, 200, 1, 5),
                                                      ('group3', 12, 60, 96, 32, 3, 9)])
def test_cnn_matches_oracle_other_shapes(method, T, E, Fk, A, window, n):
    from digat_b200 import synth
    from digat_b200.newsEncoders import CNN
    cfg = synth.make_text_config(vocabulary_size=300, max_title_length=T, word_embedding_dim=E, attention_dim=A,
                                 cnn_method=method, cnn_kernel_num=Fk, cnn_window_size=window)
    sd = synth.make_cnn_state_dict(cfg, seed=T)
    tok, mask = synth.make_titles(cfg, n, seed=T + 1)
    m = CNN(cfg)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    with torch.no_grad():
        got = m(tok.cuda().view(1, n, T), mask.cuda().view(1, n, T)).cpu()
        ref = O.cnn_news_encoder(O.cast_params(sd, torch.float64), tok.view(1, n, T), mask.double().view(1, n, T), method)
    assert rel_err(got.numpy(), ref.numpy()) < 1e-5
    bad = tok.clone().view(1, n, T)
    bad[0, 0, 0] = 300                                       # == vocabulary_size: the zero-padding row is not a token
    with pytest.raises(RuntimeError):
        m(bad.cuda(), mask.cuda().view(1, n, T))


def _grad_check(enc, ref_fn, sd, tok, mask, tol=2e-5):
    """Training forward + backward of a news encoder (p = 0, train mode) against the fp64 autograd of the oracle restatement.
    Gate per tensor: tol, or four times the error of the oracle's own fp32 autograd where that is larger -- the gradients of
    the pooling attention (affine1 / affine2) are sums over the tokens of terms whose softmax factor sums to zero, so on
    attention-smoothed token rows (MSA) they cancel to ~1e-4 of their terms in ANY fp32 arithmetic."""
    g = torch.Generator().manual_seed(3)
    w = None
    grads = {}
    for dt in (torch.float64, torch.float32):
        P = {k: v.to(dt).clone().requires_grad_(True) for k, v in sd.items()}
        ref = ref_fn(P, dt)
        if w is None:
            w = torch.randn(ref.shape, generator=g, dtype=torch.float64)
            ref64 = ref.detach()
        (ref * w.to(dt)).sum().backward()
        grads[dt] = {k: v.grad for k, v in P.items()}
    out = enc(tok.cuda(), mask.cuda())
    (out * w.float().cuda()).sum().backward()
    torch.cuda.synchronize()
    assert rel_err(out.detach().cpu().numpy(), ref64.numpy()) < 1e-5
    worst, bad = 0.0, []
    for k, v in enc.named_parameters():
        assert v.grad is not None, 'no gradient for ' + k
        e = rel_err(v.grad.cpu().numpy(), grads[torch.float64][k].numpy())
        eref = rel_err(grads[torch.float32][k].numpy(), grads[torch.float64][k].numpy())
        worst = max(worst, e)
        print('   grad %-40s ours vs fp64 %.2e   oracle fp32 vs fp64 %.2e' % (k, e, eref))
        bad = bad + [(k, e, eref)] if e >= max(tol, 4 * eref) else bad
    assert not bad, bad
    return worst


@pytest.mark.parametrize('T,heads,dk,E,A,n', [(32, 16, 25, 300, 256, 48), (20, 4, 16, 100, 64, 33), (7, 2, 32, 52, 200, 5)])
def test_msa_training_gradients_match_oracle(T, heads, dk, E, A, n):
    from digat_b200 import synth
    from digat_b200.newsEncoders import MSA
    cfg = synth.make_text_config(vocabulary_size=300, max_title_length=T, word_embedding_dim=E, MSA_head_num=heads,
                                 MSA_head_dim=dk, attention_dim=A, dropout_rate=0.0)
    sd = synth.make_msa_state_dict(cfg, seed=T)
    tok, mask = synth.make_titles(cfg, n, seed=T + 1)
    enc = MSA(cfg)
    enc.load_state_dict(sd)
    enc = enc.cuda().train()
    worst = _grad_check(enc, lambda P, dt: O.msa_news_encoder(P, tok.view(1, n, T), mask.to(dt).view(1, n, T), heads, dk), sd,
                        tok.view(1, n, T), mask.view(1, n, T))
    print('MSA T=%d: worst gradient rel err vs fp64 %.3e' % (T, worst))


@pytest.mark.parametrize('method,T,E,Fk,A,window,n', [('naive', 32, 300, 400, 256, 3, 24), ('naive', 20, 100, 64, 64, 5, 33),
                                                      ('group3', 12, 60, 96, 32, 3, 9)])
def test_cnn_training_gradients_match_oracle(method, T, E, Fk, A, window, n):
    from digat_b200 import synth
    from digat_b200.newsEncoders import CNN
    cfg = synth.make_text_config(vocabulary_size=300, max_title_length=T, word_embedding_dim=E, attention_dim=A,
                                 cnn_method=method, cnn_kernel_num=Fk, cnn_window_size=window, dropout_rate=0.0)
    sd = synth.make_cnn_state_dict(cfg, seed=T)
    tok, mask = synth.make_titles(cfg, n, seed=T + 1)
    enc = CNN(cfg)
    enc.load_state_dict(sd)
    enc = enc.cuda().train()
    worst = _grad_check(enc, lambda P, dt: O.cnn_news_encoder(P, tok.view(1, n, T), mask.to(dt).view(1, n, T), method), sd,
                        tok.view(1, n, T), mask.view(1, n, T))
    print('CNN %s T=%d: worst gradient rel err vs fp64 %.3e' % (method, T, worst))


def test_model_trains_end_to_end_from_tokens():
    """reference Model.forward on token tensors with gradients (model.py:54-77, trainer.py:98-105): news encoder -> DIGAT ->
    logits -> loss; every parameter of both encoders receives a finite gradient, dropout on."""
    import torch.nn.functional as F
    from digat_b200 import synth
    from digat_b200.model import Model
    cfg = synth.make_text_config(graph_depth=2, dropout_rate=0.2)
    model = Model(cfg)
    model.graph_encoder.load_state_dict(synth.make_state_dict(cfg, seed=3))
    model.news_encoder.load_state_dict(synth.make_msa_state_dict(cfg, seed=4))
    model = model.cuda().train()
    corpus = synth.make_corpus(cfg, n_news=60, n_behaviors=6, mean_candidates=3.0, seed=2)
    tok, mask = synth.make_titles(cfg, 60, seed=9)
    beh = np.arange(3)
    cand = np.array([[5, 9], [11, 3], [7, 20]])
    t = torch.from_numpy
    args = (tok[corpus.history[beh]], mask[corpus.history[beh]], t(corpus.user_graph[beh]), t(corpus.user_category_mask[beh]),
            t(corpus.user_category_indices[beh]), tok[corpus.news_node_ID[cand]], mask[corpus.news_node_ID[cand]],
            t(corpus.news_graph[cand]), t(corpus.news_graph_mask[cand]))
    logits = model(*[a.cuda() for a in args])
    loss = (-F.log_softmax(logits, dim=1).select(1, 0)).mean()
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss)
    for k, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
    assert float(model.news_encoder.word_embedding.weight.grad.abs().sum()) > 0
