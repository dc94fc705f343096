"""FlatAdam (digat_grad_sumsq + digat_adam_clip_step) against torch's clip_grad_norm_ + optim.Adam, the optimizer step of
reference trainer.py:98-105, and the graphed training step on a small encoder."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

STEP_TOL = 2e-6      # parameters after 6 optimizer steps, relative to max |p| (fp32 arithmetic in a different association)


@pytest.mark.parametrize('max_norm,weight_decay,shapes', [
    (1.0, 0.0, [(400, 400), (400,), (1200, 400), (19, 400), (1, 400)]),
    (0.0, 0.0, [(33, 7), (5,), (3,)]),                    # no clipping; total size not a multiple of 4
    (0.05, 0.01, [(64, 129), (129,), (2, 3, 5)]),
    (1e6, 0.0, [(250, 40)]),                              # norm below max_norm: coefficient clamps to 1
])
def test_flat_adam_matches_torch_clip_and_adam(max_norm, weight_decay, shapes):
    from digat_b200.training import FlatAdam
    g = torch.Generator().manual_seed(5)
    ref = [torch.nn.Parameter((torch.randn(*s, generator=g) * 0.1).cuda()) for s in shapes]
    ours = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    opt = torch.optim.Adam(ref, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)
    flat = FlatAdam(ours, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay, max_norm=max_norm)
    assert all(p.data_ptr() >= flat.flat_params.data_ptr() for p in ours)
    for step in range(6):
        grads = [torch.randn(*s, generator=g).cuda() * (10.0 if step % 2 else 0.01) for s in shapes]
        for p, gr in zip(ref, grads):
            p.grad = gr.clone()
        if max_norm > 0:
            norm_ref = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        else:
            norm_ref = torch.linalg.vector_norm(torch.cat([x.reshape(-1) for x in grads]))
        opt.step()
        flat.zero_grad()
        for p, gr in zip(ours, grads):
            p.grad.add_(gr)                                # what autograd's AccumulateGrad does with the persistent views
        flat.step()
        torch.cuda.synchronize()
        assert abs(float(flat.grad_norm) - float(norm_ref)) <= 2e-6 * float(norm_ref)
        for a, b in zip(ours, ref):
            scale = float(b.detach().abs().max())
            assert float((a.detach() - b.detach()).abs().max()) <= STEP_TOL * scale, (step, tuple(a.shape))
            if max_norm > 0:
                assert torch.allclose(a.grad, b.grad, rtol=2e-6, atol=0)        # the clipped gradient is written back
    assert float(flat.step_count) == 6.0


def test_graphed_step_with_flat_adam_trains_like_eager_torch():
    """A few whole training steps (forward, backward kernels, FlatAdam) replayed from one CUDA graph follow the same
    trajectory as the eager step with torch.optim.Adam + clip_grad_norm_ (dropout off: both arms see the same numbers)."""
    import torch.nn.functional as F
    from digat_b200 import synth
    from digat_b200.model import Model
    from digat_b200.training import FlatAdam, GraphedTrainStep
    cfg = synth.make_config(SAG_neighbors=3, SAG_hops=2, graph_depth=2, dropout_rate=0.0)
    sd = synth.make_state_dict(cfg, D=400, seed=3)
    corpus = synth.make_corpus(cfg, D=400, n_news=600, n_behaviors=64, mean_candidates=6.0, seed=1)
    dev = torch.device('cuda:0')

    def build():
        m = Model(cfg, 400)
        m.graph_encoder.load_state_dict(sd)
        return m.to(dev).train()
    m_ref, m_ours = build(), build()
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(0))
    emb = torch.from_numpy(corpus.news_embeddings).to(dev)
    node = torch.from_numpy(corpus.news_node_ID.astype(np.int64)).to(dev)
    ng, nm = torch.from_numpy(corpus.news_graph).to(dev), torch.from_numpy(corpus.news_graph_mask).to(dev)
    hist = torch.from_numpy(corpus.history.astype(np.int64)).to(dev)
    ug, cm, ci = (torch.from_numpy(x).to(dev) for x in (corpus.user_graph, corpus.user_category_mask, corpus.user_category_indices))

    def inputs():
        beh = torch.from_numpy(rng.integers(0, hist.shape[0], size=6)).to(dev)
        cand = torch.from_numpy(rng.integers(1, emb.shape[0], size=(6, 5))).to(dev)
        return (emb[hist[beh]], ug[beh], cm[beh], ci[beh], emb[node[cand]], ng[cand], nm[cand])
    batches = [inputs() for _ in range(8)]

    def loss_of(m, inp):
        return (-F.log_softmax(m.forward_embeddings(*inp), dim=1).select(1, 0)).mean()

    opt = torch.optim.Adam(m_ref.parameters(), lr=1e-3)
    flat = FlatAdam(m_ours.parameters(), lr=1e-3, max_norm=1.0)

    def our_step(*inp):
        loss = loss_of(m_ours, inp)
        flat.zero_grad()
        loss.backward()
        flat.step()
        return loss
    gstep = GraphedTrainStep(our_step, batches[0], warmup=3, modules=(m_ours,))      # 3 real steps on batches[0] (a capture does not execute)
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        loss_of(m_ref, batches[0]).backward()
        torch.nn.utils.clip_grad_norm_(m_ref.parameters(), 1.0)
        opt.step()
    for inp in batches[1:]:
        l_ours = float(gstep(*inp))
        opt.zero_grad(set_to_none=True)
        l = loss_of(m_ref, inp)
        l.backward()
        torch.nn.utils.clip_grad_norm_(m_ref.parameters(), 1.0)
        opt.step()
        assert abs(l_ours - float(l)) <= 1e-3 * max(1.0, abs(float(l))), (l_ours, float(l))
    torch.cuda.synchronize()
    assert float(flat.step_count) == 3 + len(batches) - 1
    worst = 0.0
    for (k, a), (_, b) in zip(m_ours.named_parameters(), m_ref.named_parameters()):
        worst = max(worst, float((a.detach() - b.detach()).abs().max()) / max(float(b.detach().abs().max()), 1e-12))
    print('worst parameter drift after 10 steps: %.3e' % worst)
    assert worst < 1e-3      # measured 5.8e-5 (Adam turns 1e-6 gradient differences into sign-level update differences on near-zero gradients)


def test_flat_adam_step_invalidates_packed_inference_weights():
    """The update kernel writes parameters in place without bumping their version counters: FlatAdam(modules=...) drops the
    encoders' packed inference copies after every step, so a no-grad forward right after a step sees the NEW weights."""
    from oracle import digat_oracle as O
    from digat_b200 import synth
    from digat_b200.graphEncoders import DIGAT
    from digat_b200.training import FlatAdam
    from tests.test_gpu_backward import ORDER, _batch
    from tests.helpers import rel_err
    cfg = synth.make_config(graph_depth=2, dropout_rate=0.0)
    sd = synth.make_state_dict(cfg, seed=2)
    batch = _batch(cfg, 6, seed=5)
    m = DIGAT(cfg, 400)
    m.load_state_dict(sd)
    m = m.cuda().eval()                                        # stays in eval mode throughout (no train()/eval() switch to help)
    b = [batch[k].cuda() for k in ORDER]
    with torch.no_grad():
        before = m(*b)[1].clone()
    flat = FlatAdam(m.parameters(), lr=1e-2, max_norm=0.0, modules=(m,))
    flat.zero_grad()
    cn, cu = m(*b)
    (cn * cu).sum().backward()
    flat.step()
    with torch.no_grad():
        after = m(*b)[1]
    torch.cuda.synchronize()
    P = {k: v.detach().double().cpu() for k, v in m.state_dict().items()}
    bb = {k: (v.double() if v.is_floating_point() else v) for k, v in batch.items()}
    ref = O.forward(P, *[bb[k] for k in ORDER])[1]
    assert not torch.allclose(before, after)
    assert rel_err(after.cpu().numpy(), ref.numpy()) < 1e-5
